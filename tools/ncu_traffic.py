"""Summarise an `ncu --set full` capture of the conv kernels into profiles/r02_conv_dram_traffic.json.
   ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > /tmp/conv_raw.csv; python tools/ncu_traffic.py /tmp/conv_raw.csv"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v, u = float(r[ix[name]].replace(",", "")), units[ix[name]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "%": 1}.get(u, 1)
    return v * scale


out = []
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    if "conv" not in name:
        continue
    out.append(dict(kernel=name.split("(")[0].replace("void ", ""), grid=r[ix["launch__grid_size"]],
                    time_us=round(val(r, "gpu__time_duration.sum") * 1e6, 1),
                    dram_read_bytes=val(r, "dram__bytes_read.sum"), dram_write_bytes=val(r, "dram__bytes_write.sum"),
                    tensor_pipe_pct=float(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in ix else None,
                    sm_throughput_pct=float(r[ix["sm__throughput.avg.pct_of_peak_sustained_elapsed"]])))
tot = sum(o["dram_read_bytes"] + o["dram_write_bytes"] for o in out)
json.dump(dict(launches=out, avg_bytes_per_launch=tot / max(1, len(out)), note="one MNIST-UNet forward, B=4096, fp16 path"),
          open("profiles/r02_conv_dram_traffic.json", "w"), indent=1)
print(len(out), "launches, avg bytes/launch", tot / max(1, len(out)))
