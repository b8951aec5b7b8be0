"""Markdown kernel table of DESIGN.md section 4 from a bench.py JSON line:  python tools/design_table.py profiles/r2_bench_final_1gpu.json"""
import json
import sys

BOUND = {"sp.conv1ab": "tensor (78 %) / drain of the conv1a result", "sp.conv1b": "tensor (at the sustained peak)", "sp.conv1a": "HBM writes / issue",
         "lg.attn_self": "tensor 65 % / MUFU / softmax chain", "lg.attn_cross": "tensor 65 % / MUFU / softmax chain",
         "lg.ffn0": "shared-memory pipe (operand reads + TMA fill)", "lg.ffn3": "HBM (775 MB per launch: operand planes + fp32 master)",
         "lg.qk": "epilogue (rotary + head split) / HBM", "lg.ln_gelu": "HBM (620 MB per launch)", "lg.vT": "HBM / epilogue, K = 256",
         "lg.out_proj": "HBM / epilogue, K = 256", "lg.assign_reduce": "L2 / HBM (4 passes over 16.8 MB per pair)", "sp.nms": "issue / shared memory",
         "sp.conv2a": "shared-memory pipe (N = 64, single CTA)", "sp.conv2b": "tensor (CTA pairs)", "sp.conv3a": "tensor", "sp.conv3b": "tensor",
         "sp.conv4a": "tensor", "sp.conv4b": "tensor", "sp.convPa": "tensor", "sp.convDa": "tensor", "sp.convDb": "tensor (small)",
         "sp.convPb": "tensor (N = 65 padded to 128)", "sp.softmax": "HBM", "sp.select+describe": "latency / gather", "lg.sim": "tensor / HBM",
         "lg.final_proj": "tensor (small)"}
d = json.load(open(sys.argv[1]))
ceil = d["roofline"]["peak"] / 3
print(f"| kernel group | ms / step | share | algorithmic TFLOP/s | of the EXACT ceiling ({ceil:.0f}) | bound by |\n|---|---|---|---|---|---|")
for k, v in d["kernels"].items():
    t = v.get("tflops_algorithmic")
    print(f"| `{k}` | {v['ms_per_step']:.2f} | {100 * v['share']:.1f} % | {t:.0f} | {100 * t / ceil:.0f} % | {BOUND.get(k, '')} |" if t else
          f"| `{k}` | {v['ms_per_step']:.2f} | {100 * v['share']:.1f} % | — | — | {BOUND.get(k, '')} |")
