"""Site merge + VCF record assembly (second half): the oracle restatement and the product's array code
against the records written by the UNMODIFIED reference find_candidates + VCFWriter (tests/golden/make_golden_vcf.py);
on the GPU the per-record flags come from the CUDA selection kernel."""
import json
import os
import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _norm(o):
    if isinstance(o, dict):
        return {k: _norm(v) for k, v in sorted(o.items())}
    if isinstance(o, (list, tuple)):
        return [_norm(x) for x in o]
    if isinstance(o, (np.floating, float)):
        return float(o)
    if isinstance(o, np.integer):
        return int(o)
    if isinstance(o, (np.bool_, bool)):
        return bool(o)
    return o


def _load():
    g = np.load(os.path.join(GOLD, "vcf_seed23.npz"))
    return g, str(g["genome"]), json.loads(str(g["options"])), [json.loads(r) for r in g["records"]]


def _fetch(genome):
    return lambda c, a, b: genome[max(0, a):max(0, b)]


def test_oracle_matches_reference_records():
    from oracle import find_candidates as ofc, vcf_records as ovr
    g, genome, options, want = _load()
    dv, off = [], 0
    for nb in g["batch_sizes"]:
        sl = slice(off, off + int(nb)); off += int(nb)
        m, d = ofc.select(options, "ctg", g["positions"][sl], g["depths"][sl], [str(k) for k in g["keys"][sl]], g["freqs"][sl], g["probs"][sl], _fetch(genome))
        dv.extend(d)
    contigs, sites = ovr.merge_sites(dv)
    got = [_norm(dict(r, id=None)) for r in ovr.vcf_records(sites, options)]
    assert got == [_norm(r) for r in want]
    assert sum(1 for r in want if len(r["alleles"]) > 2) > 50 and any(r["filter"] == "refCall" for r in want)


def _flags_from_oracle(g, genome, options):
    """Per-record flags / genotypes as the CUDA selection kernel defines them, derived from the oracle's tuples (CPU test)."""
    from oracle import find_candidates as ofc
    from pepper_b200.candidates import F_DV, F_REPEAT, F_SWAP
    n = len(g["positions"])
    flags = np.zeros(n, np.uint8); geno = np.zeros(n, np.uint8)
    for i in range(n):
        m, d = ofc.select(options, "ctg", g["positions"][i:i + 1], g["depths"][i:i + 1], [str(g["keys"][i])], g["freqs"][i:i + 1], g["probs"][i:i + 1],
                          _fetch(genome))
        geno[i] = int(np.argmax(g["probs"][i]))
        if d:
            key = str(g["keys"][i])
            swapped = key[0] == "3" and d[0][3] == key[1:]
            flags[i] = F_DV | (F_REPEAT if d[0][11] else 0) | (F_SWAP if swapped else 0)
    return flags, geno


def _product_inputs(g, genome):
    from pepper_b200 import synth
    n = len(g["positions"])
    keys = np.zeros((n, 64), np.uint8)
    for i, k in enumerate(g["keys"]):
        b = str(k).encode(); keys[i, :len(b)] = np.frombuffer(b, np.uint8)
    ref = np.frombuffer(genome.encode(), np.uint8).copy()
    regions = synth.RegionTable(np.array([[0, len(genome) - 1, 0, len(genome) - 1, 0, len(genome), 0, 0]], np.int64), ref)
    return keys, regions


def test_product_assembly_matches_reference_records():
    from pepper_b200.vcf import assemble_sites, format_vcf_line
    g, genome, options, want = _load()
    keys, regions = _product_inputs(g, genome)
    flags, geno = _flags_from_oracle(g, genome, options)
    n = len(g["positions"])
    got = assemble_sites("ctg", g["positions"], np.zeros(n, np.int32), g["depths"], g["freqs"], keys, g["probs"], flags, geno, regions, options)
    assert [_norm(dict(r, id=None)) for r in got] == [_norm(r) for r in want]
    line = format_vcf_line(got[0])
    assert line.split("\t")[0] == "ctg" and int(line.split("\t")[1]) == got[0]["start"] + 1 and line.count("\t") == 9
    assert assemble_sites("ctg", g["positions"], np.zeros(n, np.int32), g["depths"], g["freqs"], keys, g["probs"], np.zeros(n, np.uint8), geno, regions, options) == []


@pytest.mark.gpu
def test_cuda_flags_plus_assembly_match_reference_records():
    from pepper_b200.vcf import find_site_records
    g, genome, options, want = _load()
    keys, regions = _product_inputs(g, genome)
    n = len(g["positions"])
    sel = {k: options[k] for k in ("snp_p_value", "insert_p_value", "delete_p_value", "snp_p_value_in_lc", "insert_p_value_in_lc",
                                   "delete_p_value_in_lc", "report_snp_above_freq", "report_indel_above_freq")}
    got = find_site_records("ctg", g["positions"], np.zeros(n, np.int32), g["depths"], g["freqs"], keys, g["probs"].astype(np.float32), regions, sel, options)
    assert [_norm(dict(r, id=None)) for r in got] == [_norm(r) for r in want]
