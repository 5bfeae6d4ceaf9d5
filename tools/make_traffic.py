#!/usr/bin/env python
"""Writes profiles/traffic.json from the `ncu --set full` summaries of the headline step's kernels
(profiles/r02_bench_{ProxL0Box,ProxLhalfBox,IproxL0Box}.ncu_full_summary.txt, produced by tools/ncu_round2.sh) and the
hash of the CUDA sources as they are NOW: run it right after copying the summaries of a capture taken from this tree.
bench.py reports roofline.traffic only while the sources still hash to this value."""
import json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_sources_hash  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {"src_sha256": kernel_sources_hash(),
       "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, `ncu --set full --clock-control none`, one launch "
               "of bench.py (tools/ncu_round2.sh); bench.py reports them only while the CUDA sources hash to src_sha256"}
for op, stem in (("prox_l0box", "ProxL0Box"), ("prox_lhalfbox", "ProxLhalfBox"), ("iprox_l0box", "IproxL0Box")):
    path = os.path.join("profiles", f"r02_bench_{stem}.ncu_full_summary.txt")
    txt = open(os.path.join(ROOT, path)).read()
    vals = {}
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(re.escape(key) + r" = ([0-9.]+) (\w+)", txt)
        vals[key] = float(m.group(1)) * UNIT[m.group(2)]
    out[op] = {"log2n": 28, "dram_bytes_per_launch": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
               "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
               "source": path}
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
