"""How fast would K1 gather if every gathered row came from L2?  (Design probe for a column-blocked K1)

Takes the user-row block of the 1 B-interaction benchmark graph (rows gather ITEM rows: a 512 MB table, 4x the L2) and
times K1's plain SpMM on it three ways:
  full     the block as it is                                   (the traffic mix of today's kernel)
  window   column ids folded into a window of W item rows        (same rows, same degrees, same random pattern, but the
           gathered table is W x 256 B: L2-resident for W <= ~300 K) -> the gather-phase rate a column-blocked pass
           could reach (excluding the read-modify-write of the partial sums)
One JSON line per case: ms, ps per nnz, "algorithmic" GB/s (264 B per nnz + 776 B per row, SURVEY §8 d).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=10_000_000)
    ap.add_argument("--items", type=int, default=2_000_000)
    ap.add_argument("--edges", type=int, default=1_000_000_000)
    ap.add_argument("--windows", default="65536,262144,1048576")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import __graft_entry__ as G
    G.build()
    import tagrec_b200 as T
    from tagrec_b200.adj import CsrGraph, spmm_raw
    from tagrec_b200.distributed import slice_csr
    dev = torch.device("cuda:0")
    U, I = args.users, args.items
    r, c = T.data.synth_bipartite_device(U, I, args.edges, dev, seed=2020)
    full = T.build_csr(U, I, (r, c), "bi_norm", dev)
    del r, c
    rp, col, val = slice_csr(full.rowptr, full.col, full.val, 0, U)
    n, num_list, norm = full.n, full.num_list, full.norm_type
    del full
    torch.cuda.empty_cache()
    x = torch.randn(n, 64, device=dev)
    y = torch.empty_like(x)
    nnz = int(rp[-1])

    def run(name, cols):
        blk = CsrGraph(n, rp, cols, val, None, None, norm, num_list, row_offset=0)
        for _ in range(2):
            spmm_raw(blk, x, out=y)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            spmm_raw(blk, x, out=y)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        print(json.dumps({"case": name, "rows": U, "nnz": nnz, "ms": ms, "ps_per_nnz": ms * 1e9 / nnz,
                          "algorithmic_GBs": (nnz * 264.0 + U * 776.0) / (ms * 1e-3) / 1e9}), flush=True)

    run("full (512 MB item table)", col)
    for w in [int(v) for v in args.windows.split(",")]:
        folded = (U + (col - U) % w).to(col.dtype).contiguous()        # columns stay sorted per row? not needed by K1
        run(f"window {w} rows ({w * 256 / 2 ** 20:.0f} MB)", folded)
        del folded


if __name__ == "__main__":
    main()
