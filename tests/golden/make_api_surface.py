"""Writes tests/golden/api_surface.json: the public/protected method names (with arity) of every class of
the reference, read from /root/reference in the build container.  The file travels; the reference does not.
Run: python tests/golden/make_api_surface.py"""
import glob
import json
import os
import re

REF = "/root/reference/src/main/java/joezie/fora_neo4j/"
SIG = re.compile(r"^\s*(public|protected)\s+(?:static\s+)?([\w<>\[\], ?.]+?)\s+(\w+)\s*\(([^)]*)\)\s*(?:throws [\w, ]+)?\s*[{;]", re.M)


def arity(args):
    depth, n = 0, 1 if args.strip() else 0
    for ch in args:
        depth += ch == "<"
        depth -= ch == ">"
        n += ch == "," and depth == 0
    return n


def main():
    out = {}
    for f in sorted(glob.glob(REF + "*.java")):
        cls = os.path.basename(f)[:-5]
        seen = []
        for m in SIG.finditer(open(f).read()):
            name = m.group(3)
            if name == "compare":           # anonymous Comparator bodies, not class methods
                continue
            item = [name, arity(m.group(4)), "%s:%d" % (os.path.basename(f), open(f).read()[:m.start(3)].count("\n") + 1)]
            if item[:2] not in [s[:2] for s in seen]:
                seen.append(item)
        out[cls] = seen
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_surface.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(dst, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
