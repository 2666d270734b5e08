#!/usr/bin/env python3
"""Build tests/golden/catt_fixture.tar.gz from the read-only reference checkout.

The bundle holds DATA only (no reference source code):
  * testSample.fq                - the reference's bundled smoke-test input (README.md:175-184)
  * aha.TRB.CDR3.CATT.csv        - the reference's own output for it (the only end-to-end golden)
  * prob.csv                     - positional AA table read at catt.jl:656
  * resource/**                  - the IMGT V/J FASTA files named in reference.jl:5-126 plus their
                                   bwa-0.7.17 .pac/.ann/.amb companions (2-bit text, offsets, N holes)

Run in the build container (where /root/reference exists):  python tests/golden/make_fixtures.py
The GPU box has no /root/reference; tests and bench.py unpack this bundle instead.
"""
import io, os, sys, tarfile

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FASTAS = [
    "resource/TR/hs/TRB/TRBV-gai6-DNA.fa", "resource/TR/hs/TRB/TRBJ-gai3-DNA.fa",
    "resource/TR/hs/TRA/TRAV-gai3-DNA.fa", "resource/TR/hs/TRA/TRAJ-gai2-DNA.fa",
    "resource/IG/hs/IGHV-gai2-DNA.fa", "resource/IG/hs/IGHJ.fa",
    "resource/TR/hs/TRGV.fa", "resource/TR/hs/TRGJ.fa",
    "resource/TR/hs/TRDV.fa", "resource/TR/hs/TRDJ.fa",
    "resource/TR/ms/TRBV-CDR3.fa", "resource/TR/ms/TRBJ-CDR3.fa",
    "resource/TR/pig/TRBV-gai1-DNA.fa", "resource/TR/pig/TRBJ-gai1-DNA.fa",
]
FILES = ["testSample.fq", "aha.TRB.CDR3.CATT.csv", "prob.csv"]
for f in FASTAS:
    FILES.append(f)
    for ext in (".pac", ".ann", ".amb"):
        FILES.append(f + ext)

def main():
    out = os.path.join(HERE, "catt_fixture.tar.gz")
    with tarfile.open(out, "w:gz", compresslevel=9) as tar:
        for rel in FILES:
            src = os.path.join(REF, rel)
            ti = tar.gettarinfo(src, arcname=rel)
            ti.mtime = 0; ti.uid = ti.gid = 0; ti.uname = ti.gname = ""
            with open(src, "rb") as fh:
                tar.addfile(ti, fh)
    print(out, os.path.getsize(out), "bytes,", len(FILES), "files")

if __name__ == "__main__":
    sys.exit(main())
