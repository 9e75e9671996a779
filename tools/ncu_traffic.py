"""Per-launch DRAM traffic / pipe utilisation of the fused kernel from an ncu report -> profiles/rNN_ncu_traffic.json

    ncu --set full --clock-control none --import-source on -k regex:k_edge5 -s 8 -c 2 -o gpurun_out/r02_edge5f \
        python tools/fwd_probe.py cfg4 4096 auto 2          (GNB_CUDA_GRAPH=0: eager launches)
    python tools/ncu_traffic.py gpurun_out/r02_edge5f.ncu-rep profiles/r02_ncu_traffic.json

The first captured launch of a core is the edge instance (2 097 152 rows), the second the node instance (262 144 rows)."""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_units=True):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]]
    if scale_units:
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6,
              "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
    return v


res = {"source": "ncu --set full --clock-control none --import-source on -k regex:k_edge5 -s 8 -c 2 python tools/fwd_probe.py cfg4 4096 auto 2 "
                 "(GNB_CUDA_GRAPH=0; B200); per launch; written by tools/ncu_traffic.py", "kernels": {}}
names = ["tc_edge_core", "tc_node_core"]
for r, name in zip(rows[2:], names):
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    res["kernels"][name] = {
        "kernel": r[ix["Kernel Name"]],
        "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
        "duration_s_under_ncu": val(r, "gpu__time_duration.sum"),
        "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False)
        if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in ix else None,
        "issue_active_pct": 100.0 * val(r, "smsp__issue_active.avg.per_cycle_active", False),
        "l2_sectors": val(r, "lts__t_sectors_srcunit_tex.sum", False),
        "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct", False),
        "l2_read_hit_pct": val(r, "lts__t_sector_op_read_hit_rate.pct", False),
        "icache_hit_pct": val(r, "sm__icc_request_hit_rate.pct", False) if "sm__icc_request_hit_rate.pct" in ix else None,
        "lsu_wavefronts_per_sm": val(r, "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", False) if "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg" in ix else None,
        "stall_long_scoreboard_per_issue": val(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", False),
        "stall_no_instruction_per_issue": val(r, "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", False),
        "registers_per_thread": int(val(r, "launch__registers_per_thread", False)),
        "block": int(val(r, "launch__block_size", False)), "grid": int(val(r, "launch__grid_size", False)),
    }
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
