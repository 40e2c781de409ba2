"""Parse the QMP_API prototypes out of the .cu sources -> (name, return type, [(ctype, argname)]).
Used by build.py to check that include/qmp_b200.h, the ctypes table in _lib.py and the sources agree."""
import glob
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))


def prototypes():
    out = []
    for path in sorted(glob.glob(os.path.join(HERE, "*.cu"))):
        src = open(path).read()
        for m in re.finditer(r"QMP_API\s+([\w\s\*]+?)\s*\b(qmp_\w+)\s*\(([^)]*)\)\s*\{", src):
            ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
            parsed = []
            for a in [x.strip() for x in args.replace("\n", " ").split(",") if x.strip() and x.strip() != "void"]:
                mm = re.match(r"(.*?)(\w+)$", a)
                parsed.append((mm.group(1).strip(), mm.group(2)))
            out.append((os.path.basename(path), name, ret, parsed))
    return out


def code_of(ctype):
    t = ctype.replace("const", "").strip()
    if "*" in t:
        return "p"
    return {"int": "i", "long long": "l", "float": "f", "double": "d", "unsigned long long": "u"}[t]


if __name__ == "__main__":
    for f, name, ret, args in prototypes():
        print(f"{name:32s} {ret:12s} {''.join(code_of(t) for t, _ in args)}   # {f}")
