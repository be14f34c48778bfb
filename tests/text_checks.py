"""Checks of the device VCF text writer (povu_b200/csrc/pv_text.cuh), shared by the CPU
(host rehearsal) and GPU test files.  Checker: the compiled reference (oracle/_ref) and the C library's number formatting."""
import os

import pytest

import call_checks as cc
from povu_b200 import api, synth


def check_af_and_integers(pv):
    # every AF = ac / an with 0 <= ac <= an <= 700 (covers 464 haplotypes) incl. the exact binary ties 0.25 / 0.75
    assert pv.selftest_text(700) == 0


def _rename(g, names):
    out = dict(g)
    out["paths"] = [(names[i] if i < len(names) else n, st, sr) for i, (n, st, sr) in enumerate(g["paths"])]
    return out


def check_shared_slots_and_prefix_files(pv, tmp):
    """Two paths of one component that share a genotype slot (same sample and haplotype id), a haplotype id outside the
    column's ploidy range, raw (non-PanSN) names, several ref paths (record order needs the device sort), per-prefix
    output (`-o dir`: one text per prefix) -- all against the compiled reference."""
    if not os.path.exists(cc.REF):
        pytest.skip("oracle/_ref/povu_ref not built")
    gfa = os.path.join(tmp, "t.gfa")
    cases = [
        (["S001#1#a", "S001#1#b", "S002#1#a", "S002#2#a", "S003#2#x", "rawname"], ["S001"]),
        (["S001#1#a", "S001#2#a", "S001#1#b", "S002#1#a", "S002#1#b", "S003#1#c"], ["S001", "S002"]),
        (["ref", "alt1", "alt2", "S009#1#z", "S009#2#z", "S009#1#y"], ["ref"]),
    ]
    for k, (names, prefixes) in enumerate(cases):
        g = _rename(synth.snp_indel_chain(40 + 7 * k, n_paths=6, seed=60 + k), names)
        synth.write_gfa(g, gfa)
        for chunk in (100, 3):
            got, st, s = cc.gfa2vcf(pv, gfa, prefixes, chunk)
            want = cc.reference_vcf(gfa, prefixes, chunk)
            assert cc.split(got, False) == cc.split(want, False), (k, chunk)
            # per-prefix files: header of that prefix + only the records of its ref paths
            for p in prefixes:
                one = s.vcf_text(chunk_size=chunk, prefix=p)
                ref_one = cc.reference_vcf(gfa, [p], chunk)
                hdr, rec = cc.split(one, False)
                assert hdr == cc.split(ref_one, False)[0]
                tags = {l.split("\t")[0] for l in rec}
                assert rec == [l for l in cc.split(got, False)[1] if l.split("\t")[0] in tags]
                assert all(t.startswith(p) for t in tags)


def check_fd_write(pv, tmp):
    g = synth.snp_indel_chain(50, n_paths=4, seed=3)
    gfa = os.path.join(tmp, "w.gfa")
    synth.write_gfa(g, gfa)
    got, _, s = cc.gfa2vcf(pv, gfa, ["S001"])
    fp = os.path.join(tmp, "w.vcf")
    n = s.vcf_write(fp)
    assert open(fp).read() == got and n == len(got)
