#!/usr/bin/env python
"""Stage isolation of eval_tc2_kernel (the CTA-pair evaluation kernel).  Builds timing-only variants of the library from
PATCHED COPIES of the sources (the product sources carry no experiment branches) into build/variants/lib_t2x<N>.so;
run each with  TAGREC_LIB=build/variants/lib_t2x<N>.so python tools/eval_bench.py --paths tf32 .
Variants 1-4 produce wrong results on purpose.

  1  no candidates (the filter never fires): MMA + TMA + TMEM drain + max-tree
  2  1 + no TMEM drain: MMA + TMA + barrier hand-offs only
  3  2 + one tcgen05.mma per tile instead of eight: TMA stream + hand-offs
  4  1 + the drain without the max-tree/compare (tcgen05.ld only)
  5  product kernel + per-warp cycle accounting printed by block (0,0) and block (1,0) (results unchanged)
  6  5 without the ring-space check of the drain rounds; 7  6 without the entry store (bisects the cost of a round)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "tag-aware-recommendation_b200", "csrc")
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402


def patch(src, n):
    def rep(old, new, count=1):
        nonlocal src
        assert src.count(old) >= 1, old
        src = src.replace(old, new) if count == 0 else src.replace(old, new, count)
    if n in (1, 2, 3, 4):
        rep("                if (m > thr_lo) {\n", "                if (false) {\n")
    if n in (2, 3):
        rep("                tmem_ld32(taddr + c * 32, v0);\n                tmem_ld32(taddr + (c + 1) * 32, v1);\n                tmem_ld_wait();\n",
            "                for (int j = 0; j < 32; ++j) { v0[j] = 0; v1[j] = 0; }\n")
    if n == 3:
        rep("for (int kk = 0; kk < 8; ++kk) {     // K = 8 per instruction", "for (int kk = 0; kk < 1; ++kk) {     // K = 8 per instruction")
    if n == 4:
        rep("                cm[c] = scan32(v0);\n                cm[c + 1] = scan32(v1);\n",
            "                cm[c] = (v0[0] == 0x7fc00001u) | (v0[31] == 0x7fc00001u); cm[c + 1] = (v1[0] == 0x7fc00001u) | (v1[31] == 0x7fc00001u);\n")
    if n in (6, 7):
        # 6: no ring-space check in the rounds; 7: additionally no entry store (results wrong)
        rep("                if (tail + n - head_seen > (uint32_t)T2_QCAP) {", "                if (false) {")
    if n == 7:
        rep("                    my_ring[seq % T2_QCAP] = ", "                    if (seq == 0xffffffffu) my_ring[seq % T2_QCAP] = ")
    if n in (5, 6, 7):
        # drain warps: cycles waiting for the accumulator / draining / in the candidate rounds / blocked on a full ring
        rep("        uint32_t tail = 0, head_seen = 0;          // warp-uniform\n",
            "        uint32_t tail = 0, head_seen = 0;          // warp-uniform\n"
            "        long long c_wait = 0, c_drain = 0, c_rounds = 0, c_full = 0, n_cand = 0, n_rounds = 0, c_r0 = 0, n_t0 = 0, c_w64 = 0, c_r64 = 0, n_r64 = 0;\n")
        rep("            mbar_wait(smem_u32(accfull + x), (t >> 1) & 1);\n            tc_fence_after();\n            const uint32_t taddr",
            "            long long t0 = clock64();\n            mbar_wait(smem_u32(accfull + x), (t >> 1) & 1);\n            tc_fence_after();\n"
            "            long long t1 = clock64(); c_wait += t1 - t0; if (t >= 64) c_w64 += t1 - t0; const long long nr_before = n_rounds;\n            const uint32_t taddr")
        rep("            if (lane == 0) mbar_arrive_leader(smem_u32(accfree + x));\n            // ---- candidates -> this warp's ring",
            "            if (lane == 0) mbar_arrive_leader(smem_u32(accfree + x));\n            long long t2 = clock64(); c_drain += t2 - t1;\n            // ---- candidates -> this warp's ring")
        rep("                    uint32_t spins = 0;\n                    do {\n                        head_seen = *my_head;",
            "                    uint32_t spins = 0;\n                    long long tf = clock64();\n                    do {\n                        head_seen = *my_head;")
        rep("                    } while (tail + n - head_seen > (uint32_t)T2_QCAP);\n",
            "                    } while (tail + n - head_seen > (uint32_t)T2_QCAP);\n                    c_full += clock64() - tf;\n")
        rep("                tail += n;\n            }\n        }\n",
            "                tail += n; n_cand += n; ++n_rounds;\n            }\n            { long long dt = clock64() - t2; c_rounds += dt; if (n_rounds == nr_before) { c_r0 += dt; ++n_t0; } else if (t >= 64) { c_r64 += dt; n_r64 += n_rounds - nr_before; } }\n        }\n"
            "        if (lane == 0 && blockIdx.x < 2 && blockIdx.y == 0)\n"
            "            printf(\"t2 block %d drain warp %d: tiles %d wait %lld drain %lld rounds %lld (ring full %lld) per tile; candidates %lld rounds %lld; empty tiles %lld at %lld cyc; after tile 64: wait %lld per tile, %lld rounds at %lld cyc\\n\", blockIdx.x, dw, n_tiles,"
            " c_wait / n_tiles, c_drain / n_tiles, c_rounds / n_tiles, c_full / n_tiles, n_cand, n_rounds, n_t0, c_r0 / (n_t0 ? n_t0 : 1), c_w64 / (n_tiles - 64), n_r64, c_r64 / (n_r64 ? n_r64 : 1));\n")
        # scorer warps: cycles in process(), batches, entries
        rep("        int idle = 0;\n        uint32_t guard = 0;\n", "        int idle = 0;\n        uint32_t guard = 0;\n        long long c_proc = 0, n_batch = 0, n_ent = 0; const long long ts0 = clock64();\n")
        rep("                    process(st[w], dw, head[w], avail, ent);\n",
            "                    { long long tp = clock64(); process(st[w], dw, head[w], avail, ent); c_proc += clock64() - tp; ++n_batch; n_ent += avail; }\n")
        rep("#pragma unroll\n        for (int w = 0; w < 2; ++w) {\n            const int dw = 2 * sj + w;\n            const int row = ((dw + 2) & 3) * 32 + lane;\n            const int ch = dw >> 2;",
            "        if (lane == 0 && blockIdx.x < 2 && blockIdx.y == 0)\n"
            "            printf(\"t2 block %d scorer %d: total %lld cycles, in process %lld, batches %lld entries %lld\\n\", blockIdx.x, sj, clock64() - ts0, c_proc, n_batch, n_ent);\n"
            "#pragma unroll\n        for (int w = 0; w < 2; ++w) {\n            const int dw = 2 * sj + w;\n            const int row = ((dw + 2) & 3) * 32 + lane;\n            const int ch = dw >> 2;")
    return src


def main():
    out = os.path.join(ROOT, "build", "variants")
    os.makedirs(out, exist_ok=True)
    for n in [int(x) for x in (sys.argv[1:] or ["1", "2", "3", "4"])]:
        tmp = f"/tmp/t2x{n}"
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.copytree(os.path.join(ROOT, "tag-aware-recommendation_b200"), os.path.join(tmp, "pkg"), ignore=shutil.ignore_patterns("*.so"))
        shutil.copytree(os.path.join(ROOT, "include"), os.path.join(tmp, "include"))
        # the sources include "../../include/tagrec_b200.h" relative to csrc
        c = os.path.join(tmp, "pkg", "csrc")
        p = os.path.join(c, "eval_tc2.cu")
        text = patch(open(p).read(), n)
        open(p, "w").write(text)
        so = os.path.join(out, f"lib_t2x{n}.so")
        cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "--threads", "0", "-Xcompiler", "-fPIC", "-shared"] + G.SOURCES + ["-o", so]
        subprocess.run(cmd, cwd=c, check=True)
        print(so)


if __name__ == "__main__":
    main()
