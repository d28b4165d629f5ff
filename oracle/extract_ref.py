"""Build-time extraction of the reference's OWN per-ray routines for oracle/_ref (TEST INFRASTRUCTURE, never committed).

The reference cannot be built here (no MPI / libxml2 / configure-generated headers) and Ray.cc / RMCRTCommon.cc do
not compile stand-alone because their headers pull in the scheduler.  Their per-ray arithmetic, however, depends on nothing but a few
value types.  This script cuts the function definitions listed below out of the reference sources WHERE THEY LIE (by signature and
brace matching, so it follows the tree rather than hard-coded line numbers) and writes them, unmodified, to
oracle/_ref/ref_rmcrt_extracted.inc (git-ignored).  oracle/ref_rmcrt_shim.cpp declares the two classes with just the members these
functions touch, includes that file and exports a C interface, so tests can compare the oracle's restatement with the reference's own
code bit for bit.  Nothing from /root/reference is copied into the repository.

    python oracle/extract_ref.py /root/reference/src oracle/_ref/ref_rmcrt_extracted.inc
"""
import re
import sys

RMCRT = "CCA/Components/Models/Radiation/RMCRT/"
WANT = [
    (RMCRT + "RMCRTCommon.cc", r"^RMCRTCommon::raySignStep\("),
    (RMCRT + "RMCRTCommon.cc", r"^RMCRTCommon::findRayDirection\("),
    (RMCRT + "RMCRTCommon.cc", r"^RMCRTCommon::ray_Origin\("),
    (RMCRT + "RMCRTCommon.cc", r"^RMCRTCommon::reflect\("),
    (RMCRT + "RMCRTCommon.cc", r"^RMCRTCommon::updateSumI \("),
    (RMCRT + "Ray.cc", r"^Ray::containsCell\("),
    (RMCRT + "Ray.cc", r"^\s*void Ray::updateSumI_ML \("),
    # the shells around the per-ray routines: boundary-flux helpers and the Data-Onion extents
    (RMCRT + "Ray.cc", r"^Ray::computeExtents\("),
    (RMCRT + "Ray.cc", r"^void Ray::rayDirection_cellFace\("),
    (RMCRT + "Ray.cc", r"^void Ray::rayLocation_cellFace\("),
    (RMCRT + "Ray.cc", r"^Ray::has_a_boundary\("),
]
MACROS = ["DEBUG", "FIXED_RAY_DIR", "SIGN", "FUZZ"]


def cut(lines, pat):
    """the definition whose signature line matches pat: from the start of its declaration (template / return type lines above) to the
    brace that closes its body"""
    idx = [i for i, l in enumerate(lines) if re.search(pat, l)]
    if len(idx) != 1:
        raise SystemExit(f"extract_ref: {pat!r} matched {len(idx)} lines")
    i = idx[0]
    start = i
    while start > 0 and lines[start - 1].strip() and not lines[start - 1].lstrip().startswith("//") and not lines[start - 1].rstrip().endswith(("}", ";")):
        start -= 1          # return type / template<> / inline lines directly above
    depth, seen, j = 0, False, i
    in_block_comment = False
    while True:
        l = lines[j]
        k = 0
        while k < len(l):
            two = l[k:k + 2]
            if in_block_comment:
                if two == "*/":
                    in_block_comment = False
                    k += 1
            elif two == "//":
                break
            elif two == "/*":
                in_block_comment = True
                k += 1
            elif l[k] == '"':
                k += 1
                while k < len(l) and l[k] != '"':
                    k += 2 if l[k] == "\\" else 1
            elif l[k] == "'":
                k += 1
                while k < len(l) and l[k] != "'":
                    k += 2 if l[k] == "\\" else 1
            elif l[k] == "{":
                depth += 1
                seen = True
            elif l[k] == "}":
                depth -= 1
            k += 1
        if seen and depth == 0:
            return start, j
        j += 1


def main(ref, out):
    chunks = []
    cache = {}
    for rel, pat in WANT:
        if rel not in cache:
            cache[rel] = open(f"{ref}/{rel}", encoding="latin-1").read().split("\n")
        lines = cache[rel]
        a, b = cut(lines, pat)
        chunks.append(f"// ---- {rel}:{a + 1}-{b + 1}\n" + "\n".join(lines[a:b + 1]) + "\n")
    # the face tables the constructor fills (Ray.cc:112-125), as the body of a member function of the look-alike class
    tables = [l for l in cache[RMCRT + "Ray.cc"] if re.match(r"\s*d_dir(IndexOrder|SignSwap)\[\w+\]\s*=", l)]
    assert len(tables) == 12, tables
    chunks.append("// ---- " + RMCRT + "Ray.cc: d_dirIndexOrder / d_dirSignSwap assignments of Ray::Ray\nvoid Ray::initFaceTables()\n{\n" + "\n".join(tables) + "\n}\n")
    macros = []
    for l in cache[RMCRT + "RMCRTCommon.cc"]:
        m = re.match(r"#define (\w+)\s", l)
        if m and m.group(1) in MACROS:
            macros.append(l)
    assert len(macros) == len(MACROS), macros
    with open(out, "w") as f:
        f.write("// GENERATED at build time by oracle/extract_ref.py from the reference tree; NOT part of the repository.\n")
        f.write("\n".join(macros) + "\n\n" + "\n".join(chunks))
    print(f"extract_ref: wrote {out} ({sum(c.count(chr(10)) for c in chunks)} lines from {len(cache)} reference files)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
