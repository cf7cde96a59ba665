//! Drop-in prover path for `ark-sumcheck` over BLS12-381 Fr, backed by `libsumcheck_b200.so`.
//!
//! SOURCE ONLY — this crate is not compiled or tested in the repository's image (no rustc/cargo there).  It shows
//! the binding a maintainer adds: the `extern "C"` block mirrors `include/sumcheck_b200.h` one to one, and the
//! wrappers keep the reference's names, argument meaning and panics
//! (`MLSumcheck::prove`, `IPForMLSumcheck::{prover_init, prove_round}`, `GKRRoundSumcheck::prove`).
//!
//! The reference crate is `#![forbid(unsafe_code)]`, so the FFI lives here, in a sibling crate.
#![allow(clippy::missing_safety_doc)]

use ark_ff::Field;
use ark_poly::DenseMultilinearExtension;
use ark_sumcheck::ml_sumcheck::data_structures::ListOfProductsOfPolynomials;
use ark_test_curves::bls12_381::Fr;
use std::os::raw::{c_char, c_int, c_void};

/// `Fr` is `Fp<MontBackend<FrConfig,4>,4>`: a newtype over `BigInt<4>([u64; 4])` in Montgomery form.  The C ABI takes
/// exactly those 4 limbs, so slices of `Fr` are passed as `*const u64` without conversion.
const _: () = assert!(core::mem::size_of::<Fr>() == 32);

#[repr(C)]
pub struct ScBlake2b512Rng {
    pub h: [u64; 8],
    pub t: [u64; 2],
    pub buf: [u8; 128],
    pub buflen: u64,
}

#[repr(C)]
pub struct ScProver {
    _private: [u8; 0],
}
#[repr(C)]
pub struct ScComm {
    _private: [u8; 0],
}

pub const SC_OK: c_int = 0;
pub const SC_ERR_PANIC_CONSTANT: c_int = -1;
pub const SC_ERR_PANIC_FIRST_ROUND_MSG: c_int = -2;
pub const SC_ERR_PANIC_MISSING_MSG: c_int = -3;
pub const SC_ERR_PANIC_NOT_ACTIVE: c_int = -4;

extern "C" {
    pub fn sc_last_error() -> *const c_char;
    pub fn sc_device_count() -> c_int;
    pub fn sc_rng_setup(rng: *mut ScBlake2b512Rng);
    pub fn sc_rng_feed_bytes(rng: *mut ScBlake2b512Rng, b: *const u8, n: usize);
    pub fn sc_rng_fill_bytes(rng: *mut ScBlake2b512Rng, dest: *mut u8, n: usize);
    pub fn sc_rng_next_u64(rng: *mut ScBlake2b512Rng) -> u64;
    pub fn sc_rng_sample_fr(rng: *mut ScBlake2b512Rng, out: *mut u64);
    pub fn sc_prover_create(
        out: *mut *mut ScProver, nv: u32, n_tables: u32, tables: *const *const u64, n_products: u32,
        coeffs: *const u64, offsets: *const u32, indices: *const u32, device: c_int,
    ) -> c_int;
    pub fn sc_prover_create_device(
        out: *mut *mut ScProver, nv: u32, n_tables: u32, d_tables: *const *const u64, n_products: u32,
        coeffs: *const u64, offsets: *const u32, indices: *const u32, device: c_int,
    ) -> c_int;
    pub fn sc_prover_destroy(p: *mut ScProver);
    pub fn sc_prover_reset(p: *mut ScProver) -> c_int;
    pub fn sc_prover_load_tables(p: *mut ScProver, tables: *const *const u64) -> c_int;
    pub fn sc_prover_set_stream(p: *mut ScProver, cuda_stream: *mut c_void) -> c_int;
    pub fn sc_prove_round(p: *mut ScProver, r_or_null: *const u64, evals_out: *mut u64) -> c_int;
    pub fn sc_prover_max_multiplicands(p: *const ScProver) -> u32;
    pub fn sc_prover_num_vars(p: *const ScProver) -> u32;
    pub fn sc_prover_round(p: *const ScProver) -> u32;
    pub fn sc_prover_randomness(p: *const ScProver, out: *mut u64, cap: u32) -> u32;
    pub fn sc_prover_push_randomness(p: *mut ScProver, r: *const u64) -> c_int;
    pub fn sc_prover_table(p: *const ScProver, j: u32, out: *mut u64, cap_elems: u64, len_out: *mut u64) -> c_int;
    pub fn sc_ml_prove(p: *mut ScProver, rng: *mut ScBlake2b512Rng, evals_out: *mut u64, randomness_out: *mut u64) -> c_int;
    pub fn sc_ml_prove_oneshot(
        nv: u32, n_tables: u32, tables: *const *const u64, n_products: u32, coeffs: *const u64, offsets: *const u32,
        indices: *const u32, device: c_int, evals_out: *mut u64, randomness_out: *mut u64,
    ) -> c_int;
    pub fn sc_serialize_proof(evals: *const u64, nv: u32, d: u32, out: *mut u8) -> usize;
    pub fn sc_prover_launch_count(p: *const ScProver) -> u64;
    pub fn sc_prover_tc_round_count(p: *const ScProver) -> u64;
    pub fn sc_release_cached_memory();
    /// `interpolate_uni_poly` (verifier.rs:139): host-side, no GPU needed
    pub fn sc_fr_interpolate(evals: *const u64, n_evals: u32, r: *const u64, out: *mut u64) -> c_int;
    pub fn sc_gkr_initialize_phase_one(
        dim: u32, nnz: u64, f1_idx: *const u64, f1_val: *const u64, f3: *const u64, g: *const u64, device: c_int,
        h_g_out: *mut u64, f1g_idx_out: *mut u64, f1g_val_out: *mut u64, nnz_g_out: *mut u64,
    ) -> c_int;
    pub fn sc_gkr_initialize_phase_two(
        dim: u32, nnz_g: u64, f1g_idx: *const u64, f1g_val: *const u64, u: *const u64, device: c_int, f1_gu_out: *mut u64,
    ) -> c_int;
    pub fn sc_gkr_start_phase1_sumcheck(out: *mut *mut ScProver, dim: u32, h_g: *const u64, f2: *const u64, device: c_int) -> c_int;
    pub fn sc_gkr_start_phase2_sumcheck(
        out: *mut *mut ScProver, dim: u32, f1_gu: *const u64, f3: *const u64, f2_u: *const u64, device: c_int,
    ) -> c_int;
    pub fn sc_gkr_prove(
        rng: *mut ScBlake2b512Rng, dim: u32, nnz: u64, f1_idx: *const u64, f1_val: *const u64, f2: *const u64,
        f3: *const u64, g: *const u64, device: c_int, phase1_out: *mut u64, phase2_out: *mut u64, u_out: *mut u64,
        v_out: *mut u64,
    ) -> c_int;
    pub fn sc_comm_get_unique_id(id_out: *mut u8) -> c_int;
    pub fn sc_comm_create(out: *mut *mut ScComm, id: *const u8, rank: c_int, n_ranks: c_int, device: c_int) -> c_int;
    pub fn sc_comm_destroy(c: *mut ScComm);
    pub fn sc_prover_create_sharded(
        out: *mut *mut ScProver, comm: *mut ScComm, nv: u32, n_tables: u32, shard_tables: *const *const u64,
        n_products: u32, coeffs: *const u64, offsets: *const u32, indices: *const u32,
    ) -> c_int;
}

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(sc_last_error()).to_string_lossy().into_owned() }
}

/// Codes -1..-4 are the reference's own panics (prover.rs:50-52, 79-81, 90-92, 96-98); anything else is a
/// device failure and maps to `Error::OtherError` where the signature returns `Result`.
fn check(rc: c_int) -> Result<(), ark_sumcheck::Error> {
    match rc {
        SC_OK => Ok(()),
        -4..=-1 => panic!("{}", last_error()),
        _ => Err(ark_sumcheck::Error::OtherError(last_error())),
    }
}

fn limbs(x: &Fr) -> *const u64 {
    x as *const Fr as *const u64
}

/// `ProverState` of the reference (prover.rs:19-33), resident in HBM.  `randomness`, `round`, `num_vars`,
/// `max_multiplicands` and the current tables are materialised on request.
pub struct ProverState {
    h: *mut ScProver,
}
impl Drop for ProverState {
    fn drop(&mut self) {
        unsafe { sc_prover_destroy(self.h) }
    }
}
impl ProverState {
    pub fn round(&self) -> usize { unsafe { sc_prover_round(self.h) as usize } }
    pub fn num_vars(&self) -> usize { unsafe { sc_prover_num_vars(self.h) as usize } }
    pub fn max_multiplicands(&self) -> usize { unsafe { sc_prover_max_multiplicands(self.h) as usize } }
    pub fn randomness(&self) -> Vec<Fr> {
        let n = unsafe { sc_prover_randomness(self.h, core::ptr::null_mut(), 0) };
        let mut v = vec![Fr::from(0u64); n as usize];
        unsafe { sc_prover_randomness(self.h, v.as_mut_ptr() as *mut u64, n) };
        v
    }
    pub fn flattened_ml_extension(&self, j: usize) -> DenseMultilinearExtension<Fr> {
        let mut len = 0u64;
        check(unsafe { sc_prover_table(self.h, j as u32, core::ptr::null_mut(), 0, &mut len) }).unwrap();
        let mut v = vec![Fr::from(0u64); len as usize];
        check(unsafe { sc_prover_table(self.h, j as u32, v.as_mut_ptr() as *mut u64, len, &mut len) }).unwrap();
        DenseMultilinearExtension::from_evaluations_vec(len.trailing_zeros() as usize, v)
    }
}

/// Flattens `ListOfProductsOfPolynomials` (data_structures.rs:25-35) to what crosses the ABI.
struct Csr {
    tables: Vec<*const u64>,
    coeffs: Vec<Fr>,
    offsets: Vec<u32>,
    indices: Vec<u32>,
}
fn flatten(poly: &ListOfProductsOfPolynomials<Fr>) -> Csr {
    let tables = poly.flattened_ml_extensions.iter().map(|t| t.evaluations.as_ptr() as *const u64).collect();
    let (mut coeffs, mut offsets, mut indices) = (Vec::new(), vec![0u32], Vec::new());
    for (c, ix) in &poly.products {
        coeffs.push(*c);
        indices.extend(ix.iter().map(|&i| i as u32));
        offsets.push(indices.len() as u32);
    }
    Csr { tables, coeffs, offsets, indices }
}

pub struct IPForMLSumcheck;
impl IPForMLSumcheck {
    /// prover.rs:49-69 — panics on `num_variables == 0` like the reference.
    pub fn prover_init(polynomial: &ListOfProductsOfPolynomials<Fr>) -> ProverState {
        let c = flatten(polynomial);
        let mut h = core::ptr::null_mut();
        check(unsafe {
            sc_prover_create(&mut h, polynomial.num_variables as u32, c.tables.len() as u32, c.tables.as_ptr(),
                             c.coeffs.len() as u32, c.coeffs.as_ptr() as *const u64, c.offsets.as_ptr(), c.indices.as_ptr(), 0)
        }).unwrap();
        ProverState { h }
    }
    /// prover.rs:74-153 — `v_msg` is the verifier's challenge (None in the first round).  Returns P(0..d).
    pub fn prove_round(state: &mut ProverState, v_msg: &Option<Fr>) -> Vec<Fr> {
        let mut out = vec![Fr::from(0u64); state.max_multiplicands() + 1];
        let r = v_msg.as_ref().map_or(core::ptr::null(), limbs);
        check(unsafe { sc_prove_round(state.h, r, out.as_mut_ptr() as *mut u64) }).unwrap();
        out
    }
}

/// The wrapper crate's concrete `Blake2b512Rng` (rng.rs:22-81): plain data shared with the library.
pub struct Blake2b512Rng(pub ScBlake2b512Rng);
impl Blake2b512Rng {
    pub fn setup() -> Self {
        let mut s = core::mem::MaybeUninit::<ScBlake2b512Rng>::uninit();
        unsafe { sc_rng_setup(s.as_mut_ptr()); Blake2b512Rng(s.assume_init()) }
    }
    pub fn feed<M: ark_serialize::CanonicalSerialize>(&mut self, msg: &M) -> Result<(), ark_sumcheck::Error> {
        let mut buf = Vec::new();
        msg.serialize_uncompressed(&mut buf)?;
        unsafe { sc_rng_feed_bytes(&mut self.0, buf.as_ptr(), buf.len()) };
        Ok(())
    }
    pub fn sample_round(&mut self) -> Fr {
        let mut r = Fr::from(0u64);
        unsafe { sc_rng_sample_fr(&mut self.0, &mut r as *mut Fr as *mut u64) };
        r
    }
}

pub struct MLSumcheck;
impl MLSumcheck {
    /// mod.rs:42-45
    pub fn prove(polynomial: &ListOfProductsOfPolynomials<Fr>) -> Result<Vec<Vec<Fr>>, ark_sumcheck::Error> {
        let mut rng = Blake2b512Rng::setup();
        Self::prove_as_subprotocol(&mut rng, polynomial).map(|r| r.0)
    }
    /// mod.rs:50-70 with the concrete transcript: one FFI call for the whole proof.
    pub fn prove_as_subprotocol(
        fs_rng: &mut Blake2b512Rng, polynomial: &ListOfProductsOfPolynomials<Fr>,
    ) -> Result<(Vec<Vec<Fr>>, ProverState), ark_sumcheck::Error> {
        let state = IPForMLSumcheck::prover_init(polynomial);
        let (nv, d) = (polynomial.num_variables, state.max_multiplicands());
        let mut flat = vec![Fr::from(0u64); nv * (d + 1)];
        check(unsafe { sc_ml_prove(state.h, &mut fs_rng.0, flat.as_mut_ptr() as *mut u64, core::ptr::null_mut()) })?;
        Ok((flat.chunks(d + 1).map(|c| c.to_vec()).collect(), state))
    }
}
