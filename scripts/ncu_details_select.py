"""Cut the `ncu --page details` dump of scripts/ncu_r2.sh down to the first launch of the kernels DESIGN.md discusses
(profiles/r2_ncu_details_selected.txt).     python scripts/ncu_details_select.py gpurun_out/r2_ncu_details.txt"""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["conv_line_tma_kernel<3, 3>", "conv_line_tma_kernel<1, 4>", "wgrad_line_tma_kernel<3, 3>", "wgrad_line_tma_kernel<1, 13>",
        "gemm_tma_kernel<128, 0>", "gemm_tma_kernel<128, 1>", "gemm_tma_kernel<128, 2>", "wgrad_gemm_tma_kernel<64>", "wgrad_reduce_batch_kernel",
        "bn_act2_fwd_kernel", "bn_act2_bwd_fused_kernel", "dice_multi_fwd_kernel", "dice_multi_bwd_kernel", "ln_metapool_fwd_kernel",
        "breg_lap_g_kernel", "breg_lap_bwd_kernel", "fp_accum_kernel", "gate_fuse_kernel<0>"]
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2_ncu_details.txt")
lines = open(src, errors="replace").read().splitlines()
starts = [i for i, l in enumerate(lines) if re.match(r"^  \S.*\(\d+, \d+, \d+\)x\(\d+, \d+, \d+\), Context", l)]
starts.append(len(lines))
seen, out = set(), []
for a, b in zip(starts, starts[1:]):
    head = lines[a].strip()
    for w in WANT:
        if w in head and w not in seen:
            seen.add(w)
            out += lines[a:b]
            break
open(os.path.join(ROOT, "profiles", "r2_ncu_details_selected.txt"), "w").write("\n".join(out) + "\n")
print("kept", len(seen), "kernels,", len(out), "lines; missing:", [w for w in WANT if w not in seen])
