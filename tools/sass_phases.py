#!/usr/bin/env python
"""Per-phase instruction budget of the fused ensemble kernel from its SASS (no GPU needed).

Compiles ONE instantiation of pr_ensemble_kernel for sm_100a with -lineinfo, disassembles it with the inline chains
(`nvdisasm -gi`) and attributes every instruction to a phase of the Newton iteration through the `// PHASE: name`
markers in pr_ensemble_kernel.cuh (a phase runs from its marker to the next one; code inlined from pr_device.cuh
belongs to the phase of its call site).  The kernel's loops are fully unrolled, so static counts of the main loop body
ARE the per-lane counts of one Newton iteration; the level refresh runs once per accepted level.

    python tools/sass_phases.py [--G 32 --M 4 --W 16 --curv 0 --rm 1 --exact 1] [--iters-per-level 17.38] [--nodes 121]

Prints a table: per phase, warp instructions per Newton iteration and per node-iteration (x 32 lanes / nodes),
split into FP64 / MUFU / SHFL / LDS+STS / other.
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.environ.get("PR_CSRC") or os.path.join(REPO, "flow_sim_b200", "csrc")    # PR_CSRC: an older source tree
KERNEL = os.path.join(CSRC, "pr_ensemble_kernel.cuh")


def phases_from_markers():
    marks = []
    for i, line in enumerate(open(KERNEL), 1):
        m = re.search(r"//\s*PHASE:\s*([\w /+-]+?)\s*(\(|$)", line)
        if m:
            marks.append((i, m.group(1).strip()))
    return marks


def classify(op):
    op = op.split(".")[0]
    if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        return "fp64"
    if op == "MUFU":
        return "mufu"
    if op == "SHFL":
        return "shfl"
    if op in ("LDS", "STS", "LDSM"):
        return "lds_sts"
    if op in ("LDG", "STG", "LD", "ST", "LDL", "STL", "LDC", "LDCU", "ATOMG", "RED", "ATOM"):
        return "mem"
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--G", type=int, default=32)
    ap.add_argument("--M", type=int, default=4)
    ap.add_argument("--W", type=int, default=16)
    ap.add_argument("--curv", type=int, default=0)
    ap.add_argument("--rm", type=int, default=1)
    ap.add_argument("--exact", type=int, default=1)
    ap.add_argument("--iters-per-level", type=float, default=17.38)
    ap.add_argument("--nodes", type=int, default=121)
    ap.add_argument("--keep", default=None, help="directory to keep the .cubin / .sass in")
    ap.add_argument("--nvcc-flags", default="")
    ap.add_argument("--ncu-csv", default=None, help="`ncu -i rep --page source --csv` of a launch of the SAME build: "
                    "dynamic per-instruction counts replace the static weights")
    ap.add_argument("--node-iterations", type=float, default=0.0, help="sum of Newton iterations x nodes of that launch")
    a = ap.parse_args()
    tb = lambda v: "true" if v else "false"
    inst = f"pr_ensemble_kernel<{a.G}, {a.M}, {a.W}, {tb(a.curv)}, {a.rm}, {tb(a.exact)}, false, false>"
    work = a.keep or tempfile.mkdtemp(prefix="sassph")
    os.makedirs(work, exist_ok=True)
    cu = os.path.join(work, "k.cu")
    open(cu, "w").write('#include "pr_ensemble_kernel.cuh"\nnamespace pr {\ntemplate __global__ void '
                        + inst + "(const __grid_constant__ DevParams);\n}\n")
    cubin = os.path.join(work, "k.cubin")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
           "-Xptxas", "-v", "-I" + CSRC, "-cubin", "-o", cubin, cu] + a.nvcc_flags.split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr)
    regs = re.findall(r"Used (\d+) registers", r.stderr)
    spills = re.findall(r"(\d+) bytes spill stores, (\d+) bytes spill loads", r.stderr)
    sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    if a.keep:
        open(os.path.join(work, "k.sass"), "w").write(sass)

    marks = phases_from_markers()
    if not marks:
        sys.exit("no // PHASE: markers in " + KERNEL)

    def phase_of(line):
        name = "prologue"
        for ln, nm in marks:
            if line >= ln:
                name = nm
        return name

    # refresh lambda body: from the 'PHASE: level refresh' marker to its END marker
    counts = collections.defaultdict(lambda: collections.Counter())
    dyn = None
    if a.ncu_csv:
        import csv
        rows = list(csv.reader(open(a.ncu_csv)))
        hdr = rows[1]
        iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        dyn = [(r[iS].strip(), int(r[iI] or 0), int(r[iN] or 0)) for r in rows[2:] if len(r) >= len(hdr) and r[0].startswith("0x")]
    dcounts = collections.defaultdict(lambda: collections.Counter())
    samples = collections.Counter()
    idx = 0
    chain = []
    in_kernel = False
    loc_re = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    ins_re = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    pending = []
    for line in sass.splitlines():
        if line.startswith("\t.section\t.text."):
            in_kernel = "pr_ensemble_kernel" in line
            continue
        if not in_kernel:
            continue
        m = loc_re.search(line)
        if m:
            if not pending or pending[-1] is None:
                pending = []
            pending.append((m.group(1), int(m.group(2))))
            if m.group(3):
                pending.append((m.group(3), int(m.group(4))))
            continue
        m = ins_re.match(line)
        if m:
            if pending and pending[-1] is not None:      # an instruction without location lines inherits the chain
                chain = list(pending)
                pending = [None]
            # phase: the level-refresh lambda wins when any frame lies inside it; else the outermost kernel-file frame
            klines = [ln for f, ln in chain if f.endswith("pr_ensemble_kernel.cuh")]
            ph = "prologue"
            if klines:
                names = [phase_of(ln) for ln in klines]
                ph = "level refresh" if "level refresh" in names else names[-1] if names[-1] != "prologue" else names[0]
            counts[ph][classify(m.group(1))] += 1
            if dyn is not None:
                src, n_exec, n_samp = dyn[idx]
                mm = re.match(r"(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", src)
                if not mm or mm.group(1).split(".")[0] != m.group(1).split(".")[0]:
                    sys.exit(f"instruction {idx}: SASS of this build ({m.group(1)}) differs from the profiled one ({src})")
                dcounts[ph][classify(m.group(1))] += n_exec
                samples[ph] += n_samp
                idx += 1
    cols = ["fp64", "mufu", "shfl", "lds_sts", "mem", "other"]
    per_iter = {"level refresh": 1.0 / a.iters_per_level}
    once = {"prologue", "epilogue"}
    scale = 32.0 / a.nodes
    print(f"{inst}: registers {regs}, spills {spills}")
    print(f"weights: main-loop phases 1 per Newton iteration, level refresh 1/{a.iters_per_level}; per node-iteration = x 32 lanes / {a.nodes} nodes")
    hdr = f"{'phase':28s}" + "".join(f"{c:>9s}" for c in cols) + f"{'total':>9s} | {'fp64/node-it':>13s}{'all/node-it':>13s}"
    print(hdr)
    tot = collections.Counter()
    tot_w = collections.Counter()
    order = ["prologue"] + [nm for _, nm in marks if nm != "END"]
    seen = []
    for ph in order:
        if ph in seen or ph not in counts:
            continue
        seen.append(ph)
        c = counts[ph]
        t = sum(c.values())
        w = 0.0 if ph in once else per_iter.get(ph, 1.0)
        print(f"{ph:28s}" + "".join(f"{c[k]:9d}" for k in cols) + f"{t:9d} | {c['fp64'] * w * scale:13.1f}{t * w * scale:13.1f}")
        for k in cols:
            tot[k] += c[k]
            tot_w[k] += c[k] * w * scale
    print(f"{'static total':28s}" + "".join(f"{tot[k]:9d}" for k in cols) + f"{sum(tot.values()):9d}")
    print(f"{'per node-iteration':28s}" + "".join(f"{tot_w[k]:9.1f}" for k in cols) + f"{sum(tot_w.values()):9.1f}")
    if dyn is not None and a.node_iterations > 0:
        print()
        print(f"DYNAMIC (ncu 'Instructions Executed' per SASS instruction; warp instructions x 32 / {a.node_iterations:.4g} "
              "node-iterations = lane slots per node-iteration; samples = share of warp stall samples)")
        print(f"{'phase':28s}" + "".join(f"{c:>9s}" for c in cols) + f"{'total':>9s}{'samples':>9s}")
        ts = sum(samples.values())
        k32 = 32.0 / a.node_iterations
        gt = collections.Counter()
        for ph in seen:
            c = dcounts[ph]
            print(f"{ph:28s}" + "".join(f"{c[k] * k32:9.1f}" for k in cols) + f"{sum(c.values()) * k32:9.1f}{samples[ph] / ts:9.3f}")
            for k in cols:
                gt[k] += c[k]
        print(f"{'total':28s}" + "".join(f"{gt[k] * k32:9.1f}" for k in cols) + f"{sum(gt.values()) * k32:9.1f}")


if __name__ == "__main__":
    main()
