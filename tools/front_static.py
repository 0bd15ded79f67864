#!/usr/bin/env python3
"""Static figures of the front kernels that bench.py puts into `roofline` - generated, never typed in.

  python tools/front_static.py sass [lib.so]                 -> SASS instruction counts of the hot loops of the CURRENT library
  python tools/front_static.py update <ncu_full_selected.csv> [lib.so]
        -> rewrites profiles/front_kernel_static.json from an `ncu --set full` capture of `python bench.py` summarised by
           tools/summarize_ncu.py (dram bytes and executed instructions per launch) plus the SASS counts of the library
           that was profiled.  tests/test_front_static.py fails when the current library's hot loop no longer has the
           instruction count recorded there, i.e. when the JSON has gone stale.

Hot loop = the largest innermost backward-branch body of the kernel; ddc_front_bt_kernel's processes one 16-sample sub-block of one
channel per lane and iteration (ddc_front.cuh: kSub), so loop instructions / 16 = executed instructions per channel-sample;
ddc_front_tc_kernel's processes one 32-sample MMA slice (the larger of its two copies, the one that masks the mixer wrap)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ua3reo-ddc-transceiver_b200", "lib", "libua3reo_b200.so")
OUT = os.path.join(ROOT, "profiles", "front_kernel_static.json")
# samples per hot-loop iteration and lane: kSub for the CUDA-core kernels, one 32-sample MMA slice for the tensor-core kernel
KERNELS = {"ddc_front_tc_kernel": 32, "ddc_front_bt_kernel": 16, "ddc_front_kernel": 16}
UNITS = 1024 * (1 << 20)                                             # channel-samples per launch of the profiled bench


def functions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, funcs = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2)))
    return funcs


def hot_loop(ins):
    """the largest INNERMOST loop: a backward-branch body that contains no other backward branch"""
    loops = []
    for addr, txt in ins:
        if "BRA" in txt:
            t = re.search(r"0x([0-9a-f]+)", txt)
            if t and int(t.group(1), 16) < addr:
                loops.append((int(t.group(1), 16), addr))
    best = None
    for tgt, addr in loops:
        if any(tgt <= a2 < addr and t2 >= tgt for t2, a2 in loops if (t2, a2) != (tgt, addr)):
            continue                                     # has a nested loop
        n = sum(1 for a, _ in ins if tgt <= a <= addr)
        if best is None or n > best[0]:
            best = (n, tgt, addr)
    if best is None:
        return 0, {}
    n, tgt, addr = best
    body = [t for a, t in ins if tgt <= a <= addr]
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
    return n, dict(hist.most_common())


def sass_counts(lib=LIB):
    res = {}
    for mangled, ins in functions(lib).items():
        for k, per in KERNELS.items():
            # plain kernels mangle as <len><name>E...; the tensor-core kernel is a template <bool STAGE> - the instantiation
            # that ships as the default is <true> (ILb1E)
            if ("%d%sE" % (len(k), k)) in mangled or ("%d%sILb1E" % (len(k), k)) in mangled:
                n, hist = hot_loop(ins)
                res[k] = {"hot_loop_sass_instructions": n, "samples_per_iteration": per,
                          "hot_loop_sass_instructions_per_unit": n / per, "hot_loop_opcodes": hist,
                          "total_sass_instructions": len(ins)}
    return res


def update(csv_path, lib=LIB):
    # tools/summarize_ncu.py layout: one row per metric, one column per profiled launch (metric, unit, launch0, launch1, ...)
    rows = {r[0]: r for r in csv.reader(open(csv_path)) if r}
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    sass = sass_counts(lib)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for c, full_name in enumerate(rows["Kernel Name"][2:], start=2):
        name = full_name.split("(")[0].split("<")[0].replace("void ", "").strip()
        if name not in KERNELS:
            continue

        def f(key, c=c):
            return float(rows[key][c].replace(",", "")) * scale.get(rows[key][1], 1.0)
        rec = data.setdefault(name, {})
        rec["traffic_bytes_per_launch"] = f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
        rec["sass_instructions_per_unit"] = f("smsp__inst_executed.sum") * 32.0 / UNITS
        if "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg" in rows:      # per-SM average x SMs = wavefronts per launch
            n_sm = f("device__attribute_multiprocessor_count")
            if rows["SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg"][c] != "no data":
                per_sm = f("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg")
            else:                                   # ncu sometimes drops the triage counter: busy fraction x elapsed cycles (1 wavefront/clk/SM)
                per_sm = f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0 * f("sm__cycles_elapsed.avg")
            rec["lsu_wavefronts_per_unit"] = per_sm * n_sm / UNITS
            rec["lsu_wavefronts_shared_per_unit"] = f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / UNITS
            rec["lsu_pipe_pct_ncu"] = f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")
        if "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed" in rows:
            rec["tensor_pipe_pct_ncu"] = f("sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed")
        rec["source"] = "ncu --set full of `python bench.py`, summarised in profiles/%s by tools/summarize_ncu.py; " \
                        "1024 channels x 2^20 samples per launch" % os.path.basename(csv_path)
    for k, v in sass.items():
        data.setdefault(k, {}).update(v)
    data["note"] = "generated by tools/front_static.py - do not edit; tests/test_front_static.py checks it against the built library"
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "update":
        update(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else LIB)
    else:
        print(json.dumps(sass_counts(sys.argv[2] if len(sys.argv) > 2 else LIB), indent=1))
