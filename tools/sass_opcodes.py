"""Per-kernel counts of the Blackwell tensor-core / TMA opcodes in the shipped library (profiles/rNN_sass_opcodes.txt):
   python tools/sass_opcodes.py [out.txt]
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTMAPF = TMA
prefetch, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.  Runs `cuobjdump -sass` on the .so (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "composable_diffusion_models_b200", "libcdm_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "ATOMG", "REDG", "RED."]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(.*", "", name)[:110]
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        for op in OPS:
            if re.search(r"\b" + re.escape(op), line):
                counts[name][op] += 1
    lines = ["# cuobjdump -sass libcdm_b200.so: opcode counts per kernel (only kernels with tensor-core / TMA / atomic opcodes)",
             "# " + " ".join(f"{o:>8s}" for o in OPS) + "  kernel"]
    tot = collections.Counter()
    for k, c in counts.items():
        tot.update(c)
        if any(c[o] for o in OPS if o not in ("SYNCS",)):
            lines.append("  " + " ".join(f"{c[o]:8d}" for o in OPS) + "  " + k)
    lines.append("  " + " ".join(f"{tot[o]:8d}" for o in OPS) + "  TOTAL (" + str(len(counts)) + " kernels)")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
