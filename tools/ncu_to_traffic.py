"""profiles/rNN_gemm_traffic.json from an ncu capture of tools/ncu_target.py gemm (feeds roofline.traffic of bench.py):
   ncu -i gpurun_out/r02_gemm.ncu-rep --page raw --csv > raw.csv; python tools/ncu_to_traffic.py raw.csv "<how it was captured>" > profiles/r02_gemm_traffic.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = lambda n: hdr.index(n)
# ncu_target.py launches every shape twice: keep the second (warm instruction cache; the data caches are flushed by ncu anyway)
shapes = [("QKV 3072->9216 + RMSNorm + RoPE", 264), ("out-proj 3072->3072 + gate*y + fp32 residual", 279),
          ("FF-up 3072->12288 + GELU", 335), ("FF-down 12288->3072 + gate*y + fp32 residual", 490)]
out = []
for i, (name, alg) in enumerate(shapes):
    r = data[2 * i + 1]
    f = lambda n: float(r[col(n)].replace(",", ""))
    out.append({"shape": name, "kernel": r[col("Kernel Name")][:40], "duration_us": f("gpu__time_duration.sum"),
                "dram_read_MB": f("dram__bytes_read.sum"), "dram_write_MB": f("dram__bytes_write.sum"), "algorithmic_MB": alg,
                "tensor_pipe_active_pct_of_active": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "tensor_pipe_active_pct_of_elapsed": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                "registers": int(f("launch__registers_per_thread")),
                "sm_ghz": f("sm__cycles_elapsed.max") / f("gpu__time_duration.sum") / 1e3})
avg = sum((o["dram_read_MB"] + o["dram_write_MB"]) for o in out) / len(out) * 1e6
print(json.dumps({"source": sys.argv[2] if len(sys.argv) > 2 else "", "gemm_kernel_avg_dram_bytes_per_launch": avg, "launches": out}, indent=1))
