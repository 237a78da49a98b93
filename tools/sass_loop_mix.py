"""Instruction mix of the main loop (largest backward branch) of a kernel of libfdtd2d<suffix>.so -- how many issue
slots of stepK_stream's row iteration are arithmetic and how many are register moves.  CPU only (cuobjdump).
    FDTD_LIB_SUFFIX=_u2 FDTD_NVCC_EXTRA=-DFDTD_K_UNROLL=2 python tools/sass_loop_mix.py 'stepK_stream<double, 2, 6>'"""
import collections, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fdtd_em_solver_b200 import _lib

want = sys.argv[1] if len(sys.argv) > 1 else "stepK_stream<double, 2, 6>"
path = _lib.build(verbose=False)
text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if cur and m:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
for mangled, name in zip(funcs, names):
    if want not in name:
        continue
    ins = funcs[mangled]
    loops = []
    for addr, s in ins:
        if "BRA" in s:
            m = re.search(r"0x([0-9a-f]+)", s)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
    lo, hi = max(loops, key=lambda x: x[1] - x[0])
    body = [s for a, s in ins if lo <= a <= hi]
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0] for s in body)
    mov = sum(n for o, n in ops.items() if o in ("MOV", "IMAD.MOV.U32", "IMAD.MOV"))
    fp = sum(n for o, n in ops.items() if o.startswith(("DADD", "DMUL", "FADD", "FMUL")))
    print("%s: %d instructions, main loop %d: %d fp, %d moves, %d shuffles, %d other" % (
        name.split("(")[0], len(ins), len(body), fp, mov, sum(n for o, n in ops.items() if o.startswith("SHFL")),
        len(body) - fp - mov - sum(n for o, n in ops.items() if o.startswith("SHFL"))))
