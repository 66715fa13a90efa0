"""Regenerate profiles/r02_imagine_traffic.json (what bench.py quotes as roofline.traffic) from an ncu report:
    python profiles/ncu_traffic.py gpurun_out/r02_pimg_v2.ncu-rep
The report is one `ncu --set full --clock-control none -k regex:imagine_persistent -c 1` capture of
`python profiles/pimg_debug.py 1024 16` (one sd_imagine_fwd call = one launch of the persistent kernel)."""
import csv
import json
import os
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda name: next((float(v), u) for h, u, v in zip(hdr, units, vals) if h == name)
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd, ru = get("dram__bytes_read.sum")
wr, wu = get("dram__bytes_write.sum")
l2, lu = get("l1tex__m_xbar2l1tex_read_bytes.sum")
out = {
    "kernel": "sd::pimg::imagine_persistent_kernel, N=1024 H=16 (one sd_imagine_fwd call)",
    "dram_bytes": rd * scale[ru] + wr * scale[wu],
    "dram_read_bytes": rd * scale[ru], "dram_write_bytes": wr * scale[wu],
    "l2_to_sm_bytes": l2 * scale[lu],
    "duration_ms_under_ncu": get("gpu__time_duration.sum")[0],
    "tensor_pipe_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")[0],
    "registers_per_thread": get("launch__registers_per_thread")[0],
    "source": os.path.basename(rep) + " (ncu --set full --clock-control none)",
}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_imagine_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
