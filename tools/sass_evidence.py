"""SASS evidence for the hot kernels (no GPU needed): plans the BASELINE workloads at full size in
dry-run mode, NVRTC-compiles their kernels for sm_100a, disassembles the cubins with cuobjdump and
writes one opcode histogram per kernel under profiles/, with the Blackwell-specific mnemonics
(UTCHMMA / LDTM = tcgen05.mma / tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = TMA bulk copy,
SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed f32x2, ST.E.*.SYS / LD.E.*.SYS = the halo
kernel's peer flag) called out.     python tools/sass_evidence.py [out_dir]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import delayrepay_b200 as dr
from delayrepay_b200 import engine, sharding
import workloads as wl

OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles")
MARK = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "ELECT", "FFMA2", "FMUL2", "FADD2", "MUFU",
        "DFMA", "LDS", "LDG", "STG", "ST", "LD", "ATOMG", "MEMBAR", "BAR", "NANOSLEEP", "REDUX", "SHFL"]


def ph(shape, dt):
    return dr.NPArray(dr.DeviceArray.empty(shape, dt))


def disasm(kern):
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
        f.write(kern.cubin)
        path = f.name
    try:
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    finally:
        os.unlink(path)
    return sass, res


def write(label, kern, note):
    sass, res = disasm(kern)
    full, short = collections.Counter(), collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            full[m.group(2)] += 1
            short[m.group(2).split(".")[0]] += 1
    lines = [f"# {label}: kernel {kern.name}", f"# {note}",
             "# static SASS opcode counts (cuobjdump -sass of the NVRTC sm_100a cubin); "
             "executed counts per unit are in the ncu summaries", ""]
    lines += [ln.strip() for ln in res.splitlines() if "REG:" in ln]
    lines.append("")
    lines.append("Blackwell / protocol mnemonics present: " + ", ".join(
        f"{k} x{short[k]}" for k in MARK if short.get(k)))
    sysops = {k: v for k, v in full.items() if ".SYS" in k or "STRONG" in k}
    if sysops:
        lines.append("system-scope / strong memory operations: " + ", ".join(f"{k} x{v}" for k, v in sorted(sysops.items())))
    lines.append("")
    for k, v in short.most_common():
        lines.append(f"{v:6d}  {k}")
    lines.append(f"{sum(short.values()):6d}  TOTAL")
    path = os.path.join(OUT, f"r2_sass_{label}.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(path, sum(short.values()), "instructions;", lines[6])


def last(prefix):
    return [k for k in engine._kernels.values() if k.name.startswith(prefix)][-1]


with engine.dry_run():
    S, K, T = (ph((1 << 30,), np.float32) for _ in range(3))
    dr.evaluate(*wl.black_scholes(dr, S, K, T))
    write("black_scholes_flat_staged", last("dr_flat_"),
          "Black-Scholes call+put, 2^30 float32 options: per-warp TMA bulk rings (UBLKCP + SYNCS), packed f32x2 arithmetic, table erf/exp/log")
    u = ph((32768, 32768), np.float32)
    wl.heat(dr, u, 1)
    write("heat_stencil", last("dr_stencil_"),
          "5-point heat stencil, 32768^2 float32: UTMALDG.2D tile ring + mbarrier, ping-pong output")
    a, b = ph((1 << 30,), np.float64), ph((1 << 30,), np.float64)
    wl.l2_distance(dr, a, b).run()
    write("l2_distance_fused_reduce", last("dr_flat_"), "sqrt(sum((a-b)**2)), 2^30 float64: fused producer + block reduction + ticketed final fold")
    pos, m = ph((65536, 3), np.float32), ph((65536,), np.float32)
    wl.nbody_acc(dr, pos, m).run()
    write("nbody_mm_skinny", last("dr_mm_skinny_"), "n-body all-pairs producer fused into W @ pos (+ row sum as a ones column), 65536 bodies")
    A, B = ph((4096, 4096), np.float32), ph((4096, 4096), np.float32)
    (A @ B).run()
    write("dense_gemm_tcgen05", last("dr_gemm") if any(k.name.startswith("dr_gemm") for k in engine._kernels.values())
          else [k for k in engine._kernels.values() if "tcgen05" in k.source or "UTCHMMA" in k.name][-1],
          "dense float32 A @ B, 4096^3: 3xTF32 tcgen05.mma with TMEM accumulator, TMA operand ring")
    mesh = sharding.init(devices=[0, 0])
    su = dr.shard(wl.make_inputs("heat", 512)["u"])
    wl.heat(dr, su, 1)
    write("heat_stencil_halo", last("dr_stencil_"),
          "row-sharded heat stencil: edge tile loop stores boundary rows into the neighbour GPU's block and publishes the step with st.release.sys; waits with ld.acquire.sys")
    sharding.shutdown()
