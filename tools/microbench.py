"""Kernel-level timings on one GPU (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sp1_b200 import Lib

def main():
    ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    log_h = int(sys.argv[2]) if len(sys.argv) > 2 else 21
    lib = Lib(0)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    msg = torch.randint(0, 0x7f000001, (ncols, 1 << log_h), dtype=torch.int32, device="cuda", generator=g)
    out = torch.empty((ncols, 1 << (log_h + 2)), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    cells = ncols << log_h
    for it in range(4):
        lib.rs_encode(msg, out, ncols, log_h, 2)
        ms = lib.phase_ms("rs_encode")
        print(f"rs_encode C={ncols} 2^{log_h}: {ms:.3f} ms  -> {20*cells/ms/1e6:.1f} GB/s algorithmic (20 B/cell), {cells/ms/1e6:.2f} Gcell/s")
    for it in range(3):
        lib.merkle_commit(out, ncols, log_h + 2)
        ms = lib.phase_ms("merkle_commit")
        perms = (1 << (log_h + 2)) * ((ncols + 7) // 8) + (1 << (log_h + 2))
        print(f"merkle_commit: {ms:.3f} ms -> {perms/ms/1e6:.2f} Gperm/s ({perms} perms)")
    n = 1 << 24
    st = torch.randint(0, 0x7f000001, (n, 16), dtype=torch.int32, device="cuda", generator=g)
    torch.cuda.synchronize()
    for it in range(3):
        lib.poseidon2_permute(st)
        ms = lib.phase_ms("poseidon2_permute")
        print(f"poseidon2_permute n=2^24: {ms:.3f} ms -> {n/ms/1e6:.2f} Gperm/s")
    print("launches", lib.launch_count())

if __name__ == "__main__":
    main()
