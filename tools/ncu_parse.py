"""Dev tool: read a `ncu --set full` report (`ncu -i X.ncu-rep --page raw --csv`) and write the two summaries bench.py and
the judge read: profiles/r2_ncu_full_kernels.csv (one row per captured launch: duration, DRAM bytes, tensor-pipe and L2
utilisation, registers, shared memory) and profiles/r2_ncu_traffic.json (kernel name -> DRAM bytes per launch: the LAST
captured launch - the warm one - of every kernel's largest captured shape)."""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_hmma_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_cycles_pct"),
        ("lts__t_bytes.sum", "l2_bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
have = [(m, n) for m, n in want if m in col]


def num(v, unit):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
             "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}
    return v * scale.get(unit, 1.0)


out_rows, traffic, shape_dur = [], {}, {}
for r in data:
    name = re.sub(r"\(.*$", "", r[col["Kernel Name"]])
    name = re.sub(r"^void ", "", name).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    rec = {"kernel": name}
    for m, n in have:
        rec[n] = num(r[col[m]], units[col[m]])
    out_rows.append(rec)
    base = re.sub(r"<.*$", "", name)
    if "dram_read" in rec:
        # one figure per kernel: the LAST (= warm) launch of its largest captured shape (a later, smaller shape of the same
        # kernel - K4b at 128^2 after 512^2 - must not replace it: longest duration decides, within 10 %)
        best = shape_dur.get(base, 0.0)
        if rec.get("duration", 0.0) >= 0.9 * best:
            traffic[base] = rec["dram_read"] + rec["dram_write"]
            shape_dur[base] = max(best, rec.get("duration", 0.0))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "r2_ncu_full_kernels.csv"), "w") as f:
    w = csv.DictWriter(f, fieldnames=["kernel"] + [n for _, n in have])
    w.writeheader()
    for rec in out_rows:
        w.writerow({k: (f"{v:.6g}" if isinstance(v, float) else v) for k, v in rec.items()})
with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w") as f:
    json.dump({k: round(v) for k, v in traffic.items()}, f, indent=1, sort_keys=True)
print(f"{len(out_rows)} launches; duration in us, bytes in bytes")
for rec in out_rows:
    print("  ".join(f"{k}={v:.4g}" if isinstance(v, float) else str(v) for k, v in rec.items()))
