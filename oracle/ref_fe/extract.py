#!/usr/bin/env python
"""Cuts the hot-path member functions out of the reference's OWN source text at build time.

TEST INFRASTRUCTURE ONLY (oracle/): nothing under nextsim_b200/ imports or links this.

    python oracle/ref_fe/extract.py /root/reference oracle/_ref/ref_fe_bodies.inc

The output (git-ignored, under oracle/_ref/) is the verbatim text of the listed function definitions of
/root/reference/model/finiteelement.cpp and core/src/gmshmesh.cpp, in file order, each preceded by a
`#line` directive so compiler diagnostics point back into the reference.  It is compiled by the Makefile of
this directory against stub_fe.hpp, a stand-in for finiteelement.hpp that declares only what those bodies
touch (Boost, Gmsh, NetCDF and MPI are absent from this image, so the real header cannot be parsed).
No reference source is copied into the repository: the .inc file is a build product.

A definition is located by its qualified name plus a distinguishing piece of its parameter list, then cut
from the first line of its declaration (return type / template header) to the brace that closes its body;
comments, string and character literals are skipped while matching braces.
"""
import os
import re
import sys

# (file, qualified name, text that must occur between the name and the opening brace, occurrence index)
WANTED = [
    ("model/finiteelement.cpp", "FiniteElement::initFETensors", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::jacobian", "std::vector<std::vector<double>> const& vertices", 0),
    ("model/finiteelement.cpp", "FiniteElement::sides", "mesh_type const& mesh) const", 0),
    ("model/finiteelement.cpp", "FiniteElement::sides", "mesh_type const& mesh,", 0),
    ("model/finiteelement.cpp", "FiniteElement::minAngles", "std::vector<double> const& um, double factor", 0),
    ("model/finiteelement.cpp", "FiniteElement::minAngle", "std::vector<double> const& um, double factor, bool root", 0),
    ("model/finiteelement.cpp", "FiniteElement::flip", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::measure", "FEMeshType const& mesh) const", 0),
    ("model/finiteelement.cpp", "FiniteElement::measure", "std::vector<double> const& um, double factor", 0),
    ("model/finiteelement.cpp", "FiniteElement::shapeCoeff", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::calcCohesion", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::update", "std::vector<double> const & UM_P", 0),
    ("model/finiteelement.cpp", "FiniteElement::updateSigmaDamage", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::updateIceDiagnostics", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::checkRegridding", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::explicitSolve", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::updateSigmaVP", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::updateSigmaEVP", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::updateSigmaMEVP", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::updateGhosts", "", 0),
    # thermo() and everything it calls (libm only)
    ("model/finiteelement.cpp", "FiniteElement::specificHumidity", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::OWBulkFluxes", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::thermo", "int dt", 0),
    ("model/finiteelement.cpp", "FiniteElement::IABulkFluxes", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::windSpeedElement", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::incomingLongwave", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::iceOceanHeatflux", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::freezingPoint", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::albedo", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::meltPonds", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::thermoWinton", "", 0),
    ("model/finiteelement.cpp", "FiniteElement::thermoIce0", "", 0),
    ("core/src/gmshmesh.cpp", "GmshMesh::vertices", "std::vector<int> const& indices) const", 0),
    ("core/src/gmshmesh.cpp", "GmshMesh::vertices", "std::vector<double> const& um, double factor", 0),
]


def skip_noncode(s, i):
    """If s[i] starts a comment or a literal, return the index just after it, else i."""
    if s.startswith("//", i):
        j = s.find("\n", i)
        return len(s) if j < 0 else j
    if s.startswith("/*", i):
        j = s.find("*/", i + 2)
        return len(s) if j < 0 else j + 2
    if s[i] in "\"'":
        q = s[i]
        j = i + 1
        while j < len(s) and s[j] != q:
            j += 2 if s[j] == "\\" else 1
        return j + 1
    return i


def find_definition(src, name, must, occurrence=0):
    pat = re.compile(r"(?<![\w:])" + re.escape(name) + r"\s*\(")
    seen = 0
    for m in pat.finditer(src):
        # walk to the opening brace of the body; a ';' first means this is a call or a declaration
        i = m.end()
        depth = 1
        while i < len(src) and depth:                      # closing parenthesis of the parameter list
            j = skip_noncode(src, i)
            if j != i:
                i = j
                continue
            depth += src[i] == "("
            depth -= src[i] == ")"
            i += 1
        j = i
        while j < len(src) and src[j] not in "{;":
            k = skip_noncode(src, j)
            j = k if k != j else j + 1
        if j >= len(src) or src[j] != "{":
            continue
        head = src[m.start():j]
        line_start = src.rfind("\n", 0, m.start()) + 1
        if src[line_start:m.start()].strip():                # something before the name on its line: a call
            continue
        if must and must not in " ".join(head.split()):
            continue
        if seen < occurrence:
            seen += 1
            continue
        # body end
        i = j + 1
        depth = 1
        while i < len(src) and depth:
            k = skip_noncode(src, i)
            if k != i:
                i = k
                continue
            depth += src[i] == "{"
            depth -= src[i] == "}"
            i += 1
        end = i
        # declaration start: the lines directly above the name that belong to it (return type, inline, template<>)
        start = line_start
        while True:
            prev_end = start - 1
            if prev_end <= 0:
                break
            prev_start = src.rfind("\n", 0, prev_end) + 1
            prev = src[prev_start:prev_end].strip()
            if not prev or prev.startswith("//") or prev.endswith("}") or prev.endswith(";") or prev.startswith("#"):
                break
            start = prev_start
        return start, end
    raise SystemExit("extract.py: definition of %s (%r) not found" % (name, must))


def main():
    ref, out = sys.argv[1], sys.argv[2]
    pieces = []
    cache = {}
    for rel, name, must, occ in WANTED:
        path = os.path.join(ref, rel)
        if path not in cache:
            cache[path] = open(path, encoding="utf-8", errors="replace").read()
        src = cache[path]
        a, b = find_definition(src, name, must, occ)
        line = src.count("\n", 0, a) + 1
        pieces.append((rel, name, line, src.count("\n", 0, b) + 1, src[a:b]))
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    with open(out, "w") as f:
        f.write("// GENERATED by oracle/ref_fe/extract.py from the reference's own sources -- build product, not tracked.\n")
        for rel, name, l0, l1, text in pieces:
            f.write("\n// ---- %s  %s:%d-%d ----\n" % (name, rel, l0, l1))
            f.write('#line %d "%s"\n' % (l0, os.path.join(ref, rel)))
            f.write(text)
            f.write("\n")
    with open(out + ".index", "w") as f:
        for rel, name, l0, l1, text in pieces:
            f.write("%s %s:%d-%d\n" % (name, rel, l0, l1))
    print("extract.py: %d definitions, %d lines -> %s" % (len(pieces), sum(p[4].count("\n") + 1 for p in pieces), out))


if __name__ == "__main__":
    main()
