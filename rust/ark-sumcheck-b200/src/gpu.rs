//! `src/gpu.rs` of the B200 fork of `ark-sumcheck` (feature `gpu`) — the ONLY module with `unsafe`.
//!
//! SOURCE ONLY: this repository's image has no rustc/cargo, so the file has never been compiled.  It is written against
//! `arkworks-rs/sumcheck` @ d241e9b and the C ABI in `include/sumcheck_b200.h`; every item cites what it binds.
//!
//! Why a fork and not a sibling crate: `ProverMsg.evaluations` is `pub(crate)` (src/ml_sumcheck/protocol/prover.rs:14-17)
//! and `GKRProof`'s fields are `pub(crate)` with no `CanonicalDeserialize` (src/gkr_round_sumcheck/data_structures.rs:9-12),
//! so code outside the crate cannot build the values the prover API returns.  Inside the crate nothing changes for callers:
//! same paths, same generic signatures, same types; `patches/` holds the six small hunks that route the prover path here
//! when `F` is BLS12-381 Fr, and keep the reference's CPU code for every other field.
#![allow(unsafe_code)]

use crate::ml_sumcheck::data_structures::ListOfProductsOfPolynomials;
use crate::ml_sumcheck::protocol::prover::ProverMsg;
use ark_ff::{BigInteger, Field, PrimeField};
use ark_poly::{DenseMultilinearExtension, SparseMultilinearExtension};
use ark_std::os::raw::{c_char, c_int};
use ark_std::string::String;
use ark_std::vec::Vec;

// ------------------------------------------------------------------------------------------------ raw ABI
/// `sc_blake2b512_rng` (include/sumcheck_b200.h): Blake2b512Rng (src/rng.rs:22-81) as plain data.
#[repr(C)]
#[derive(Clone)]
pub struct ScBlake2b512Rng {
    h: [u64; 8],
    t: [u64; 2],
    buf: [u8; 128],
    buflen: u64,
}
#[repr(C)]
pub struct ScProver {
    _private: [u8; 0],
}

extern "C" {
    fn sc_last_error() -> *const c_char;
    fn sc_device_count() -> c_int;
    fn sc_rng_setup(rng: *mut ScBlake2b512Rng);
    fn sc_rng_feed_bytes(rng: *mut ScBlake2b512Rng, b: *const u8, n: usize);
    fn sc_rng_fill_bytes(rng: *mut ScBlake2b512Rng, dest: *mut u8, n: usize);
    fn sc_prover_create(
        out: *mut *mut ScProver, nv: u32, n_tables: u32, tables: *const *const u64, n_products: u32, coeffs: *const u64,
        offsets: *const u32, indices: *const u32, device: c_int,
    ) -> c_int;
    fn sc_prover_create_multi(
        out: *mut *mut ScProver, nv: u32, n_tables: u32, tables: *const *const u64, n_products: u32, coeffs: *const u64,
        offsets: *const u32, indices: *const u32, device_ids: *const c_int, n_devices: u32,
    ) -> c_int;
    fn sc_prover_destroy(p: *mut ScProver);
    fn sc_prove_round(p: *mut ScProver, r_or_null: *const u64, evals_out: *mut u64) -> c_int;
    fn sc_prover_push_randomness(p: *mut ScProver, r: *const u64) -> c_int;
    fn sc_prover_table(p: *const ScProver, j: u32, out: *mut u64, cap_elems: u64, len_out: *mut u64) -> c_int;
    fn sc_ml_prove(p: *mut ScProver, rng: *mut ScBlake2b512Rng, evals_out: *mut u64, randomness_out: *mut u64) -> c_int;
    fn sc_gkr_initialize_phase_one(
        dim: u32, nnz: u64, f1_idx: *const u64, f1_val: *const u64, f3: *const u64, g: *const u64, device: c_int,
        h_g_out: *mut u64, f1g_idx_out: *mut u64, f1g_val_out: *mut u64, nnz_g_out: *mut u64,
    ) -> c_int;
    fn sc_gkr_initialize_phase_two(
        dim: u32, nnz_g: u64, f1g_idx: *const u64, f1g_val: *const u64, u: *const u64, device: c_int, f1_gu_out: *mut u64,
    ) -> c_int;
    fn sc_gkr_prove(
        rng: *mut ScBlake2b512Rng, dim: u32, nnz: u64, f1_idx: *const u64, f1_val: *const u64, f2: *const u64, f3: *const u64,
        g: *const u64, device: c_int, phase1_out: *mut u64, phase2_out: *mut u64, u_out: *mut u64, v_out: *mut u64,
    ) -> c_int;
}

// ------------------------------------------------------------------------------------------------ field gate
/// BLS12-381 Fr modulus, little-endian u64 limbs.
const FR_MODULUS: [u64; 4] = [0xffffffff00000001, 0x53bda402fffe5bfe, 0x3339d80809a1d805, 0x73eda753299d7d48];

/// True when `F` is a 4-limb Montgomery prime field with the BLS12-381 Fr modulus — `ark_bls12_381::Fr`,
/// `ark_test_curves::bls12_381::Fr`, or any other `Fp<MontBackend<_, 4>, 4>` over that modulus: 32 bytes, `[u64; 4]` Montgomery
/// limbs (R = 2^256), which is exactly what the C ABI takes.  Every other field keeps the reference's CPU path.
pub fn is_gpu_field<F: Field>() -> bool {
    F::extension_degree() == 1
        && core::mem::size_of::<F>() == 32
        && core::mem::align_of::<F>() == 8
        && <F::BasePrimeField as PrimeField>::MODULUS.to_bytes_le()
            == FR_MODULUS.iter().flat_map(|l| l.to_le_bytes()).collect::<Vec<u8>>()
        && device_available()
}

/// Devices to use: `SUMCHECK_B200_DEVICES=0,1,2,3` shards one proof over several GPUs from this one process
/// (sc_prover_create_multi); default device 0.  No visible CUDA device => the CPU path (the C library itself never falls back).
fn devices() -> Vec<c_int> {
    std::env::var("SUMCHECK_B200_DEVICES")
        .ok()
        .map(|s| s.split(',').filter_map(|x| x.trim().parse().ok()).collect())
        .filter(|v: &Vec<c_int>| !v.is_empty())
        .unwrap_or_else(|| vec![0])
}
fn device_available() -> bool {
    static ONCE: std::sync::OnceLock<bool> = std::sync::OnceLock::new();
    *ONCE.get_or_init(|| unsafe { sc_device_count() } > 0)
}

#[inline]
fn limbs<F: Field>(x: &[F]) -> *const u64 {
    debug_assert!(is_gpu_field::<F>());
    x.as_ptr() as *const u64
}
#[inline]
fn limbs_mut<F: Field>(x: &mut [F]) -> *mut u64 {
    x.as_mut_ptr() as *mut u64
}

fn last_error() -> String {
    unsafe { core::ffi::CStr::from_ptr(sc_last_error()) }.to_string_lossy().into_owned()
}
/// SC_ERR_PANIC_* (-1..-4) are the reference's own panics (prover.rs:50-52, 79-81, 90-92, 96-98) with the same text; every
/// other failure is a device/driver problem the reference's signatures (no `Result`) cannot carry: panic with the message.
fn check(rc: c_int) {
    if rc != 0 {
        panic!("{}", last_error());
    }
}
/// ... and where the signature does return `Result`: `Error::OtherError` (src/error.rs:19).
fn check_result(rc: c_int) -> Result<(), crate::Error> {
    match rc {
        0 => Ok(()),
        -4..=-1 => panic!("{}", last_error()),
        _ => Err(crate::Error::OtherError(last_error())),
    }
}

// ------------------------------------------------------------------------------------------------ transcript
/// The state `Blake2b512Rng` holds under feature `gpu` (patches/0003-rng.patch): the same hash chain, byte for byte
/// (csrc/blake2b.cuh is pinned against RFC 7693 vectors and an independent model), as plain data so that a whole proof can run
/// inside ONE library call with the transcript advanced in place.
pub struct GpuTranscript(pub(crate) ScBlake2b512Rng);
impl GpuTranscript {
    pub(crate) fn setup() -> Self {
        let mut st = core::mem::MaybeUninit::<ScBlake2b512Rng>::uninit();
        unsafe {
            sc_rng_setup(st.as_mut_ptr());
            GpuTranscript(st.assume_init())
        }
    }
    pub(crate) fn feed_bytes(&mut self, b: &[u8]) {
        unsafe { sc_rng_feed_bytes(&mut self.0, b.as_ptr(), b.len()) }
    }
    pub(crate) fn fill_bytes(&mut self, dest: &mut [u8]) {
        unsafe { sc_rng_fill_bytes(&mut self.0, dest.as_mut_ptr(), dest.len()) }
    }
}

// ------------------------------------------------------------------------------------------------ prover handle
/// Owner of an `sc_prover` (ProverState resident in HBM).  Lives in the private `gpu` field the fork adds to `ProverState`.
pub struct GpuProver {
    raw: *mut ScProver,
    n_tables: usize,
    d: usize,
}
impl Drop for GpuProver {
    fn drop(&mut self) {
        unsafe { sc_prover_destroy(self.raw) }
    }
}

impl GpuProver {
    /// `IPForMLSumcheck::prover_init` (prover.rs:49-69): the deep copy of every unique table becomes the H2D upload.
    pub(crate) fn new<F: Field>(poly: &ListOfProductsOfPolynomials<F>) -> Self {
        let tables: Vec<*const u64> = poly.flattened_ml_extensions.iter().map(|t| limbs(&t.evaluations)).collect();
        let coeffs: Vec<F> = poly.products.iter().map(|(c, _)| *c).collect();
        let mut offsets: Vec<u32> = Vec::with_capacity(poly.products.len() + 1);
        let mut indices: Vec<u32> = Vec::new();
        offsets.push(0);
        for (_, ix) in &poly.products {
            indices.extend(ix.iter().map(|&j| j as u32));
            offsets.push(indices.len() as u32);
        }
        let devs = devices();
        let mut raw = core::ptr::null_mut();
        let rc = unsafe {
            if devs.len() > 1 {
                sc_prover_create_multi(
                    &mut raw, poly.num_variables as u32, tables.len() as u32, tables.as_ptr(), poly.products.len() as u32,
                    limbs(&coeffs), offsets.as_ptr(), indices.as_ptr(), devs.as_ptr(), devs.len() as u32,
                )
            } else {
                sc_prover_create(
                    &mut raw, poly.num_variables as u32, tables.len() as u32, tables.as_ptr(), poly.products.len() as u32,
                    limbs(&coeffs), offsets.as_ptr(), indices.as_ptr(), devs[0],
                )
            }
        };
        check(rc); // nv == 0 -> panic!("Attempt to prove a constant.")
        GpuProver { raw, n_tables: tables.len(), d: poly.max_multiplicands }
    }

    /// `IPForMLSumcheck::prove_round` (prover.rs:74-153): fold on `r` + the d+1 sums, one call.
    pub(crate) fn prove_round<F: Field>(&mut self, r: Option<&F>) -> ProverMsg<F> {
        let mut evaluations = vec![F::zero(); self.d + 1];
        let rp = r.map_or(core::ptr::null(), |x| limbs(core::slice::from_ref(x)));
        check(unsafe { sc_prove_round(self.raw, rp, limbs_mut(&mut evaluations)) });
        ProverMsg { evaluations }
    }

    /// The whole of `MLSumcheck::prove_as_subprotocol`'s round loop (ml_sumcheck/mod.rs:54-67) for the concrete
    /// `Blake2b512Rng`: PolynomialInfo is fed, every round runs with the transcript, the last challenge is pushed.
    pub(crate) fn prove_all<F: Field>(&mut self, rng: &mut GpuTranscript, nv: usize) -> Result<(Vec<ProverMsg<F>>, Vec<F>), crate::Error> {
        let mut evals = vec![F::zero(); nv * (self.d + 1)];
        let mut randomness = vec![F::zero(); nv];
        check_result(unsafe { sc_ml_prove(self.raw, &mut rng.0, limbs_mut(&mut evals), limbs_mut(&mut randomness)) })?;
        let msgs = evals.chunks(self.d + 1).map(|c| ProverMsg { evaluations: c.to_vec() }).collect();
        Ok((msgs, randomness))
    }

    /// `prover_state.randomness.push(r)` without folding (ml_sumcheck/mod.rs:65-67), mirrored on the device-side state.
    pub(crate) fn push_randomness<F: Field>(&mut self, r: &F) {
        check(unsafe { sc_prover_push_randomness(self.raw, limbs(core::slice::from_ref(r))) });
    }

    /// `ProverState.flattened_ml_extensions` at the current round (length 2^(nv-round+1)); the reference keeps these on the
    /// host after every fold, the fork downloads them when asked (`ProverState::sync_from_device`) and at the end of
    /// `prove_as_subprotocol` (two elements per table).
    pub(crate) fn tables<F: Field>(&self) -> Vec<DenseMultilinearExtension<F>> {
        (0..self.n_tables)
            .map(|j| {
                let mut len = 0u64;
                check(unsafe { sc_prover_table(self.raw, j as u32, core::ptr::null_mut(), 0, &mut len) });
                let mut v = vec![F::zero(); len as usize];
                check(unsafe { sc_prover_table(self.raw, j as u32, limbs_mut(&mut v), len, &mut len) });
                DenseMultilinearExtension::from_evaluations_vec(len.trailing_zeros() as usize, v)
            })
            .collect()
    }
}

// ------------------------------------------------------------------------------------------------ GKR initialisers
fn flatten_sparse<F: Field>(f: &SparseMultilinearExtension<F>) -> (Vec<u64>, Vec<F>) {
    // ark-poly keeps the nonzeros in a BTreeMap<usize, F>: unique, sorted indices
    f.evaluations.iter().map(|(i, v)| (*i as u64, *v)).unzip()
}

/// `initialize_phase_one` (gkr_round_sumcheck/mod.rs:22-42).
pub(crate) fn initialize_phase_one<F: Field>(
    f1: &SparseMultilinearExtension<F>, f3: &DenseMultilinearExtension<F>, g: &[F],
) -> (DenseMultilinearExtension<F>, SparseMultilinearExtension<F>) {
    let dim = f3.num_vars;
    let (idx, val) = flatten_sparse(f1);
    let mut h_g = vec![F::zero(); 1 << dim];
    let (mut gi, mut gv) = (vec![0u64; idx.len().max(1)], vec![F::zero(); idx.len().max(1)]);
    let mut n_g = 0u64;
    check(unsafe {
        sc_gkr_initialize_phase_one(
            dim as u32, idx.len() as u64, idx.as_ptr(), limbs(&val), limbs(&f3.evaluations), limbs(g), devices()[0],
            limbs_mut(&mut h_g), gi.as_mut_ptr(), limbs_mut(&mut gv), &mut n_g,
        )
    });
    let pairs: Vec<(usize, F)> = gi[..n_g as usize].iter().zip(&gv).map(|(i, v)| (*i as usize, *v)).collect();
    (DenseMultilinearExtension::from_evaluations_vec(dim, h_g), SparseMultilinearExtension::from_evaluations(2 * dim, &pairs))
}

/// `initialize_phase_two` (mod.rs:57-63).
pub(crate) fn initialize_phase_two<F: Field>(f1_g: &SparseMultilinearExtension<F>, u: &[F]) -> DenseMultilinearExtension<F> {
    let dim = u.len();
    let (idx, val) = flatten_sparse(f1_g);
    let mut out = vec![F::zero(); 1 << dim];
    check(unsafe {
        sc_gkr_initialize_phase_two(dim as u32, idx.len() as u64, idx.as_ptr(), limbs(&val), limbs(u), devices()[0], limbs_mut(&mut out))
    });
    DenseMultilinearExtension::from_evaluations_vec(dim, out)
}

/// The whole of `GKRRoundSumcheck::prove` (mod.rs:93-139) for the concrete `Blake2b512Rng`: two phases, one call.
pub(crate) fn gkr_prove<F: Field>(
    rng: &mut GpuTranscript, f1: &SparseMultilinearExtension<F>, f2: &DenseMultilinearExtension<F>,
    f3: &DenseMultilinearExtension<F>, g: &[F],
) -> (Vec<ProverMsg<F>>, Vec<ProverMsg<F>>) {
    let dim = f2.num_vars;
    let (idx, val) = flatten_sparse(f1);
    let (mut m1, mut m2) = (vec![F::zero(); dim * 3], vec![F::zero(); dim * 3]);
    check(unsafe {
        sc_gkr_prove(
            &mut rng.0, dim as u32, idx.len() as u64, idx.as_ptr(), limbs(&val), limbs(&f2.evaluations), limbs(&f3.evaluations),
            limbs(g), devices()[0], limbs_mut(&mut m1), limbs_mut(&mut m2), core::ptr::null_mut(), core::ptr::null_mut(),
        )
    });
    let split = |m: Vec<F>| m.chunks(3).map(|c| ProverMsg { evaluations: c.to_vec() }).collect();
    (split(m1), split(m2))
}
