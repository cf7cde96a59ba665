for nt in 1 0; do for kb in 1024 2048 4096 16384; do for ns in 4 8; do for th in 8 12; do
echo -n "NT=$nt slotKB=$kb slots=$ns threads=$th: "; SC_COPY_NT=$nt SC_BOUNCE_SLOT_KB=$kb SC_BOUNCE_SLOTS=$ns SC_COPY_THREADS=$th python tools/e2e_probe.py 2>&1 | tail -1; done; done; done; done
