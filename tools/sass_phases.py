"""Static SASS accounting used in DESIGN.md 7.5-7.8 (no GPU needed): instructions between the barriers of a kernel
(`nvdisasm --print-line-info` order = program order for these straight-line phase programs) and the order of global
loads / shared stores / atomics (is a batch of loads issued before the first dependent store?).

    python tools/sass_phases.py <object.o> <kernel name substring> [...]

Build objects with python -m packnet_sfm_b200.build (packnet_sfm_b200/build/*.o).  Output of round 1:
profiles/r01c_static_sass_analysis.txt."""
import os
import re
import subprocess
import sys
import tempfile


def disasm(obj):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    return subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cubin)], check=True, stdout=subprocess.PIPE, text=True).stdout


def kernels(text):
    out, name = {}, None
    for ln in text.split("\n"):
        m = re.match(r"^(_Z\w+):$", ln)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        if ln.startswith("\t.section") or re.match(r"^//-+ \.", ln):
            name = None
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*);", ln)
        if m and name:
            out[name].append(re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip()))
    return out


def main():
    obj, pats = sys.argv[1], sys.argv[2:]
    ks = kernels(disasm(obj))
    sym = {"LDG": "L", "STG": "S", "STS": "s", "BAR": "|", "ATOMS": "a", "ATOMG": "A", "REDG": "r", "RED": "r", "SHFL": "x"}
    for name, ins in ks.items():
        if pats and not any(p in name for p in pats):
            continue
        dem = subprocess.run(["c++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip()
        print("%s\n  %d instructions" % (dem[:150], len(ins)))
        bars = [i for i, t in enumerate(ins) if t.startswith("BAR.SYNC")]
        prev = 0
        segs = []
        for b in bars + [len(ins)]:
            segs.append(b - prev)
            prev = b
        print("  instructions between barriers:", segs)
        ops = {}
        for t in ins:
            k = t.split()[0].split(".")[0]
            ops[k] = ops.get(k, 0) + 1
        print("  top opcodes:", sorted(ops.items(), key=lambda kv: -kv[1])[:12])
        print("  memory / sync order:", "".join(sym.get(t.split()[0].split(".")[0], "") for t in ins)[:220])


if __name__ == "__main__":
    main()
