"""Per-source-line profile of one kernel: joins the SASS page of an ncu report (instructions executed, stall samples
per instruction) with the line table of the same cubin (nvdisasm -g), in instruction order.

    python profiles/sass_by_line.py report.ncu-rep libmtfjsp_b200.so 'Li6ELi6ELi8ELi4EEELi11EfE' [topN]

The .so must be the build the report was captured from (the join asserts that the opcodes agree)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


HDR = {}


def sass_lines(so, pattern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    txt = []
    for cub in sorted((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: -os.path.getsize(os.path.join(tmp, f))):
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
        if any(ln.startswith("//--------------------- .text.") and pattern in ln for ln in txt):
            break  # the cubin (one per .cu file) that holds the kernel
    out, on, line = [], False, 0
    for ln in txt:
        if ln.startswith("//--------------------- .text."):
            on = pattern in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "(.*)", line (\d+)', ln)
        if m:
            # lines of other files (CUDA headers: shuffles, math) are folded into negative pseudo-lines per header
            line = int(m.group(2)) if m.group(1).endswith(".cu") else -(abs(hash(os.path.basename(m.group(1)))) % 1000 + 1)
            if line < 0:
                HDR[line] = os.path.basename(m.group(1))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), line))
    return out


def main():
    rep, so, pat = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    names = rows[h]
    ci, cs, cx = names.index("Source"), names.index("# Samples"), names.index("Instructions Executed")
    stall_cols = [(n, names.index(n)) for n in names if n.startswith("stall_") and "Not Issued" not in n]
    ins = [r for r in rows[h + 1:] if len(r) == len(names)]
    sl = sass_lines(so, pat)
    assert len(sl) >= len(ins) * 0.9, (len(sl), len(ins))
    by = {}
    tot_s = tot_x = 0
    n = min(len(sl), len(ins))
    mism = 0
    for k in range(n):
        op_a = ins[k][ci].strip().split()[0 if not ins[k][ci].strip().startswith("@") else 1].split(".")[0]
        op_b = sl[k][1].split()[0 if not sl[k][1].startswith("@") else 1].split(".")[0]
        if op_a != op_b:
            mism += 1
        line = sl[k][2]
        d = by.setdefault(line, {"samples": 0, "inst": 0, "stalls": {}})
        s, x = int(ins[k][cs] or 0), int(ins[k][cx] or 0)
        d["samples"] += s; d["inst"] += x
        tot_s += s; tot_x += x
        for nme, idx in stall_cols:
            v = int(ins[k][idx] or 0)
            if v:
                d["stalls"][nme] = d["stalls"].get(nme, 0) + v
    print("instructions %d (ncu) / %d (nvdisasm), opcode mismatches %d; samples %d, warp-instructions executed %d" % (
        len(ins), len(sl), mism, tot_s, tot_x))
    src = open(os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", "mtfjsp_env.cu")).read().splitlines()
    print("%5s %7s %6s %7s %6s  %-40s %s" % ("line", "samples", "%", "inst", "%", "top stalls", "source"))
    for line, d in sorted(by.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:3]
        print("%5d %7d %6.2f %7d %6.2f  %-40s %s" % (line, d["samples"], 100.0 * d["samples"] / max(tot_s, 1), d["inst"],
              100.0 * d["inst"] / max(tot_x, 1), " ".join("%s:%d" % (a.replace("stall_", ""), b) for a, b in st),
              src[line - 1].strip()[:90] if 0 < line <= len(src) else HDR.get(line, "")))
    # phases: contiguous line ranges
    print("\ncumulative by line range")
    # phases of env_kernel_s, located by marker strings in the source (line numbers move with the code)
    marks = [("prologue: action, staging issue", "__global__ void __launch_bounds__(S::WARPS * 32, S::MINB) env_kernel_s"),
             ("policy draw, dependent loads, barrier wait", "    if constexpr ((MODE & MODE_POLICY) != 0) {\n        // uniform random selectable job"),
             ("mfea1 (policy)", "        // candidate-machine features of the drawn op (trainer/parallel_env.py:152-214), lane k = machine k"),
             ("placement scan", "    if (MODE & MODE_STEP) {\n        const int nsched0 = s_misc[2];"),
             ("estimator chain + apply", "        // estimator chain of the job's remaining ops (SS:1964-1995): lane c ends up with op (ja, c)"),
             ("idle sum", "        // ---- idle time: sequential sum in (machine, route) order, DGenv_func.py:144-170 ----"),
             ("per-job state", "    // per job (lane j): ops scheduled so far, ESA key"),
             ("energy sum (incremental)", "        // ---- energy estimate (SS:896): np.sum over all ops"),
             ("reward + scaler + write-back", "        // ---- reward (SS:1066-1132) and reward scaling"),
             ("job mask", "    // ---- job mask + candidates (ppo_algorithm.py:202-317) ----"),
             ("observation rows", "        // feature row (SS:2246-2277) and compact ELL adjacency row (SS:2019-2073) of op v; eptv = its estimated energy"),
             ("(end)", "// ---- static tables: min feasible duration / energy per op")]
    text = "\n".join(src)
    starts = []
    for nm, mk in marks:
        k = text.rfind(mk)
        starts.append((nm, text.count("\n", 0, k) + 1 if k >= 0 else None))
    ranges = [(starts[i][1], starts[i + 1][1] - 1, starts[i][0]) for i in range(len(starts) - 1) if starts[i][1] and starts[i + 1][1]]
    kstart = starts[0][1] or 0
    for a, b, nm in ranges:
        s = sum(d["samples"] for l, d in by.items() if a <= l <= b)
        x = sum(d["inst"] for l, d in by.items() if a <= l <= b)
        print("  %4d-%4d %-45s samples %5.1f%%  inst %5.1f%%" % (a, b, nm, 100.0 * s / max(tot_s, 1), 100.0 * x / max(tot_x, 1)))
    s = sum(d["samples"] for l, d in by.items() if 0 <= l < kstart)
    x = sum(d["inst"] for l, d in by.items() if 0 <= l < kstart)
    print("  helpers (above the kernel: reductions, adj_val_t, rand)       samples %5.1f%%  inst %5.1f%%" % (100.0 * s / max(tot_s, 1), 100.0 * x / max(tot_x, 1)))
    s = sum(d["samples"] for l, d in by.items() if l < 0)
    x = sum(d["inst"] for l, d in by.items() if l < 0)
    print("  CUDA headers (shuffles, ballots, math intrinsics)            samples %5.1f%%  inst %5.1f%%" % (100.0 * s / max(tot_s, 1), 100.0 * x / max(tot_x, 1)))


if __name__ == "__main__":
    main()
