"""profiles/traffic.json from the summary of an `ncu --set full` capture of the step kernel (tools/ncu_summary.py output),
stamped with the hash of the kernel sources of THIS tree: bench.py copies roofline.traffic / ncu_pipe_utilisation_pct from
it only while the hash still matches (bench.committed_ncu_figures).

    python tools/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep profiles/<tag>_ncu_step_tc_kernel_summary.csv
    python tools/make_traffic_json.py profiles/<tag>_ncu_step_tc_kernel_summary.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

src = sys.argv[1]
rows = {r[0]: r[1:] for r in csv.reader(open(src)) if r}


def val(metric, launch=0):
    unit, v = rows[metric][0], float(rows[metric][1 + launch])
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit, 1.0)
    return v * scale


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
pct = lambda m: round(val(m), 1)
out = {
    "dram_bytes_per_launch": int(rd + wr), "read": int(rd), "write": int(wr),
    "kernel": "pinn_step_tc_kernel<2,true,false>", "points": 262144, "algorithmic_bytes": 16 * 262144,
    "source": "%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % os.path.relpath(src, ROOT),
    "kernel_source_sha256": bench.kernel_source_sha256(),
    "gpu_time_us": val("gpu__time_duration.sum"), "sm_cycles_elapsed_max": val("sm__cycles_elapsed.max"),
    "ncu_pipe_utilisation_pct": {
        "issue_active": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fma": pct("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "tensor": pct("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "alu": pct("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "lsu": pct("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "xu": pct("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "shared_memory_wavefronts": pct("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "source": "%s (ncu --set full, pct_of_peak_sustained_active)" % os.path.relpath(src, ROOT),
    },
}
out["effective_sm_clock_mhz_under_ncu"] = round(out["sm_cycles_elapsed_max"] / out["gpu_time_us"], 1)
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
