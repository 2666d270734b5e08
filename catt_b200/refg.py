"""The `refg` data contract of the reference (reference.jl:5-126): per species/chain/region the V/J FASTA
paths (relative to the CATT `resource/` directory), the Cys / Phe-Trp motif regexes with their offsets,
the inner rescue motifs and the optional D-segment list.  The GPU path consumes exactly these fields.
"""
from __future__ import annotations

import os

REFG = {
    "hs": {
        "TRA": {"CDR3": {
            "vregion": "TR/hs/TRA/TRAV-gai3-DNA.fa", "jregion": "TR/hs/TRA/TRAJ-gai2-DNA.fa",
            "cmotif": "[EHILMSTV]{1}Y[FILY]{1}C[AGILV]{1}",
            "fmotif": "(LA|YI|FI|II|LY|LM|PT|TI|LV|ST|VT|LT|LI|LQ|MR|VI|FV|FQ|LF|LL|FE|FT|LS|LN|FY)F((ARG)|(G[A-Z]{1}G))",
            "coffset": 3, "foffset": 0, "innerC": "place_holder", "innerF": "place_holder"}},
        "TRB": {"CDR3": {
            "vregion": "TR/hs/TRB/TRBV-gai6-DNA.fa", "jregion": "TR/hs/TRB/TRBJ-gai3-DNA.fa",
            "cmotif": "(LR|YF|YI|YL|YQ|YR)C(A|S|T|V|G|R|P|D)", "fmotif": "[FTYH]{1}FG[ADENPQS]{1}G",
            "coffset": 2, "foffset": 1, "innerC": "(CAS|CSA|CAW|CAT|CSV|CAI|CAR)", "innerF": "place_holder",
            "Dregion": [("TRBD1*01", "gggacagggggc"), ("TRBD2*01", "gggactagcggggggg"), ("TRBD2*02", "gggactagcgggaggg")]}},
        "TRG": {"CDR3": {
            "vregion": "TR/hs/TRGV.fa", "jregion": "TR/hs/TRGJ.fa",
            "cmotif": "(YY|YH)C(A|T)", "fmotif": "([LV]{1}FG[SP]{1}G)|(IFAEG)|(TFAKG)",
            "coffset": 2, "foffset": 1, "innerC": "(CAT|CAW|CTT|CAL|CAC|CAA)", "innerF": "(KLF)|(KVF)|(KIF)|(KTF)"}},
        "TRD": {"CDR3": {
            "vregion": "TR/hs/TRDV.fa", "jregion": "TR/hs/TRDJ.fa",
            "cmotif": "(YF|YL|YY)C(A|S)", "fmotif": "[I|F]FG[K|T]G",
            "coffset": 2, "foffset": 1, "innerC": "place_holder", "innerF": "place_holder",
            "Dregion": [("TRDD1*01", "gaaatagt"), ("TRDD2*01", "ccttcctac"), ("TRDD3*01", "actgggggatacg")]}},
        "IGH": {"CDR3": {
            "vregion": "IG/hs/IGHV-gai2-DNA.fa", "jregion": "IG/hs/IGHJ.fa",
            "cmotif": "(LYY|TAV|YVA|TYY|MYY|VYD|AFN|AYY|EYY|THY|VYY|LYH|VDY)C", "fmotif": "[VLHYPIS]{1}WG[QR]{1}G",
            "coffset": 3, "foffset": 1, "innerC": "place_holder", "innerF": "place_holder"}},
    },
    "ms": {
        "TRB": {"CDR3": {
            "vregion": "TR/ms/TRBV-CDR3.fa", "jregion": "TR/ms/TRBJ-CDR3.fa",
            "cmotif": "[AGPST]{1}C(LHV|FYI|LYF|LFV|FYV|LYL|MYL|LYT|TCY|LYV|FYL|FYT|LYM|LYI|FYM)",
            "fmotif": "(SF|FF|YF|HF|TF|DH|LF)G[ADMREKLSPH]{1}G",
            "coffset": 0, "foffset": 1, "innerC": "place_holder", "innerF": "place_holder"}},
    },
    "pig": {
        "TRB": {"CDR3": {
            "vregion": "TR/pig/TRBV-gai1-DNA.fa", "jregion": "TR/pig/TRBJ-gai1-DNA.fa",
            "cmotif": "(FF|FL|LF|YF|YL|YV)C", "fmotif": "[HILTYFM]{1}FG[ADEGLPS]{1}G",
            "coffset": 2, "foffset": 1, "innerC": "place_holder", "innerF": "(LH|LI|LT|LY|QH|QI|QY|RY|VF|YN)F",
            "Dregion": [("TRBD1*01", "gggacagggggggc"), ("TRBD2*01", "ggagctatgggggggg"), ("TRBD3*01", "ggagctatggggggggg")]}},
    },
}


def chain_config(species: str, chain: str, region: str = "CDR3", resource_dir: str | None = None) -> dict:
    """refg[species][chain][region] with FASTA paths resolved against `resource_dir` (catt.jl:710-723 raises
    for unsupported combinations; so do we)."""
    try:
        cfg = dict(REFG[species][chain][region])
    except KeyError as e:
        raise KeyError(f"Selected {species}-{chain}-{region} is not supported") from e
    if resource_dir is not None:
        cfg["vregion"] = os.path.join(resource_dir, cfg["vregion"])
        cfg["jregion"] = os.path.join(resource_dir, cfg["jregion"])
    return cfg
