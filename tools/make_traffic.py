"""profiles/<visit>/traffic.json (DRAM bytes per launch of each kernel family, read by bench.py for roofline.traffic) from an
.ncu-rep:  python tools/make_traffic.py gpurun_out/rXX/prof.ncu-rep profiles/rXX/traffic.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
fam = {"tc_forward_kernel": "field_forward", "tc_bwd2_kernel": "field_backward", "tc_bwd_kernel": "field_backward", "composite_loss_kernel": "integral_loss",
       "adam_kernel": "adam"}
def val(r, name, to):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "%": 1.0}.get(u, 1.0)
    return v * scale
per = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    key = next((f for k, f in fam.items() if k in name), None)
    if key is None or key in per:
        continue
    rd, wr = val(r, "dram__bytes_read.sum", "b"), val(r, "dram__bytes_write.sum", "b")
    per[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "time_us": val(r, "gpu__time_duration.sum", "us"),
                "tensor_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "%"),
                "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%"), "dram_bytes": rd + wr}
json.dump({"source": f"ncu --set full --clock-control none of one launch each (1024 rays x 500 samples), {rep}", "per_launch": per}, open(out, "w"), indent=1)
print(json.dumps(per, indent=1))
