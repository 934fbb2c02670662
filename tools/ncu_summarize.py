"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel family -> markdown table."""
import csv
import re
import sys
from collections import defaultdict


def family(name):
    for key, fam in (("conv_tc_persistent", "conv_tc persistent (tcgen05 implicit GEMM)"), ("conv_thin", "conv_thin (tcgen05, thin layers)"),
                     ("conv_halo_persistent", "conv_halo_persistent (tcgen05, halo reuse, persistent)"), ("conv_halo_fused", "conv_halo_fused (tcgen05, halo reuse, persistent, GroupNorm+SiLU on the operand path; width-folded thin layers, upsample phases)"), ("conv_halo", "conv_halo one-tile (tcgen05, N=16 layers)"),
                     ("conv_small", "conv_small (streaming 1x1 / stride-2 / stem)"), ("conv_tc_kernel", "conv_tc one-tile (qkv / split)"), ("attention_pipe", "attention (tcgen05 flash, pipelined S/P/O)"), ("attention_kernel", "attention (tcgen05 flash, serial / 3xTF32)"),
                     ("conv_fixed", "conv_fixed (unrolled streaming 1x1 / stride-2)"), ("gn_tile_reduce", "groupnorm fold of epilogue statistics"),
                     ("conv_direct_kernel", "conv_direct"), ("gn_partial", "groupnorm stats"), ("gn_finalize", "groupnorm finalize"),
                     ("gn_apply", "groupnorm apply+SiLU"), ("upsample", "upsample"), ("fbp_filter", "fbp filter"), ("fbp_backproject", "fbp backproject"),
                     ("moments", "sampler moments"), ("finalize", "sampler finalize"), ("apply_kernel", "sampler apply"),
                     ("select_", "median select"), ("delta_map", "delta map"), ("lambda_step", "lambda map"), ("qsample", "q_sample"),
                     ("lincomb", "lincomb"), ("clamp", "clamp"), ("sharpen", "sharpen"), ("set_int", "set t")):
        if key in name:
            return fam
    return "other: " + name[:40]


def main(path):
    rows = list(csv.reader(l for l in open(path, errors="ignore") if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}.get(r[ui], 1e-3)
        f = family(r[ki])
        agg[f] += v * scale
        cnt[f] += 1
    tot = sum(agg.values())
    print(f"| kernel family | launches | total us | share |\n|---|---:|---:|---:|")
    for f, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"| {f} | {cnt[f]} | {v:.0f} | {100 * v / tot:.1f}% |")
    print(f"| **all** | {sum(cnt.values())} | {tot:.0f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1])
