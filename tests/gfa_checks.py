"""Device GFA parser (povu_b200/csrc/pv_gfa.cuh) against its two checkers: the host loader of
pv_host.hpp (same liteseq / from_gfa.cpp:28-113 rules, sequential) and the independent Python reader of helpers.py.
Shared by the host-rehearsal test (CPU) and the GPU test."""
import os
import random

import numpy as np

import helpers
from povu_b200 import api, synth

ARRAYS = ["ids", "seg_len", "ep1", "ep2", "step_off", "steps", "hap_len", "label_off"]


def _load(pv, path, host, labels=True, refs=True):
    """-> ("ok", arrays) or ("err", message)"""
    try:
        s = api.Session(pv, path, inc_vtx_labels=labels, inc_refs=refs, host_parser=host)
    except RuntimeError as e:
        return "err", str(e)
    try:
        return "ok", s.graph_arrays(labels=labels)
    finally:
        s.close()


def same_as_host(pv, path, labels=True, refs=True):
    kd, d = _load(pv, path, False, labels, refs)
    kh, h = _load(pv, path, True, labels, refs)
    assert kd == kh, f"{path}: device parse -> {kd} ({d if kd == 'err' else ''}), host loader -> {kh} ({h if kh == 'err' else ''})"
    if kd == "err":
        assert d == h, f"{path}: messages differ\n device: {d}\n host:   {h}"
        return kd, d
    for k in ARRAYS:
        if k in h:
            assert np.array_equal(d[k], h[k]), f"{path}: array {k} differs"
    assert d["path_names"] == h["path_names"]
    if labels:
        assert d["labels"] == h["labels"]
    return kd, d


def check_fixtures(pv, tmp):
    for name in helpers.POSITIVE + helpers.SYNTHETIC + ["LPA"]:
        path = helpers.fixture_path(name, tmp)
        kind, d = same_as_host(pv, path)
        assert kind == "ok", name
        g = helpers.read_gfa(path)                       # independent reader
        ep1, ep2 = api.edge_ports(g)
        assert np.array_equal(d["ids"], g["ids"]) and np.array_equal(d["ep1"], ep1) and np.array_equal(d["ep2"], ep2), name
        assert d["labels"] == "".join(g["seqs"]).encode(), name
        assert d["path_names"] == [p[0] for p in g["paths"]], name
        for k, p in enumerate(g["paths"]):
            st = d["steps"][d["step_off"][k]:d["step_off"][k + 1]]
            assert np.array_equal(st, (p[1].astype(np.uint32) << 1) | p[2]), (name, k)
            assert d["hap_len"][k] == sum(len(g["seqs"][v]) for v in p[1]), (name, k)
    for name in helpers.NEGATIVE:
        kind, msg = same_as_host(pv, helpers.fixture_path(name, tmp))
        assert kind == "err" and "liteseq could not parse the file" in msg, (name, msg)


EDGE = {
    "unsorted_dups": "S\t7\tAC\nS\t3\tG\nS\t7\tTTTT\nS\t5\tCA\nL\t3\t+\t5\t-\t0M\nL\t7\t-\t3\t+\t*\nP\tx#1#c\t3+,5-,7+\t*\n",
    "crlf_no_final_newline": "H\tVN:Z:1.0\r\nS\t1\tACGT\r\nS\t2\tT\r\n\r\nL\t1\t+\t2\t+\t0M\r\nP\tp\t1+,2+\t*",
    "w_lines_blank_lines_tags": "H\tVN:Z:1.1\n\nS\t1\tA\tLN:i:1\tRC:i:7\nW\ts\t1\tc\t0\t9\t>1>2\nS\t2\tCC\n\nL\t1\t+\t2\t+\t0M\tXX:Z:tag\nP\tq#2#chr\t1+,2+\t0M,0M\tAB:Z:1,2,3\n",
    "sparse_ids": "S\t5\tA\nS\t4000000000\tCG\nS\t77\tT\nL\t5\t+\t4000000000\t+\t0M\nL\t4000000000\t-\t77\t-\t0M\nP\tz\t5+,4000000000+,77-\t*\n",
    "single_step_paths": "S\t1\tA\nS\t2\tC\nP\ta\t1+\t*\nP\tb\t2-\t*\nP\tc#1#k\t1-,1-,1+\t*\n",
    "no_links_no_paths": "S\t10\tACGTACGT\nS\t11\tA\n",
    "self_loops": "S\t1\tAC\nS\t2\tG\nL\t1\t+\t1\t+\t0M\nL\t1\t+\t1\t-\t0M\nL\t2\t-\t2\t-\t0M\nL\t1\t+\t2\t+\t0M\nP\tr\t1+,1+,2+\t*\n",
    "non_numeric_ids": "S\tchr1\tA\nS\t2\tC\nS\t-3\tG\nL\t2\t+\t2\t+\t0M\n",
    "odd_orientations": "S\t1\tA\nS\t2\tC\nL\t1\t?\t2\t++\t0M\nL\t1\t-\t2\t\t0M\n",
}

ERRORS = {
    "s_no_tab": "S\t1\tA\nS\n",
    "s_one_tab": "S\t1\tA\nL\t1\t+\t1\t+\t0M\nS\t2\n",
    "s_empty_seq": "S\t1\t\nS\t2\tA\n",
    "s_empty_seq_tab": "S\t1\t\tLN:i:0\n",
    "s_seq_at_end": "S\t1\tA\nS\t2\t",
    "unsupported_type": "S\t1\tA\n\n\nX\tfoo\nS\t2\n",
    "unsupported_first_of_two": "S\t1\nQ\tbar\n",
    "comment_line": "# a comment\nS\t1\tA\n",
    "no_segments": "H\tVN:Z:1.0\nL\t1\t+\t2\t+\t0M\n",
    "empty_file": "",
    "only_newlines": "\n\n\n",
    "link_overlap_1m": "S\t1\tA\nS\t2\tC\nL\t1\t+\t2\t+\t1M\n",
    "link_unknown_segment": "S\t1\tA\nS\t2\tC\nL\t1\t+\t3\t+\t0M\n",
    "link_five_fields": "S\t1\tA\nS\t2\tC\nL\t1\t+\t2\t+\n",
    "link_bad_id": "S\t1\tA\nS\t2\tC\nL\t1x\t+\t2\t+\t0M\n",
    "link_empty_overlap": "S\t1\tA\nS\t2\tC\nL\t1\t+\t2\t+\t\n",
    "path_no_overlaps_column": "S\t1\tA\nP\tp\t1+\n",
    "path_two_fields": "S\t1\tA\nP\tp\n",
    "path_only_type": "S\t1\tA\nP\n",
    "path_empty_walk": "S\t1\tA\nP\tp\t\t*\n",
    "path_bad_orientation": "S\t1\tA\nS\t2\tC\nP\tp\t1+,2\t*\n",
    "path_trailing_comma": "S\t1\tA\nP\tp\t1+,\t*\n",
    "path_leading_comma": "S\t1\tA\nP\tp\t,1+\t*\n",
    "path_unknown_segment": "S\t1\tA\nP\tp\t1+,9+\t*\n",
    "path_bare_sign": "S\t1\tA\nP\tp\t+\t*\n",
    "second_path_bad": "S\t1\tA\nP\tp\t1+\t*\nP\tq\t1+,1\t*\n",
}


def _write(tmp, name, text):
    p = os.path.join(tmp, name + ".gfa")
    with open(p, "wb") as f:
        f.write(text.encode())
    return p


def check_edge_cases(pv, tmp):
    for name, text in EDGE.items():
        kind, d = same_as_host(pv, _write(tmp, name, text))
        assert kind == "ok", (name, d)
        same_as_host(pv, _write(tmp, name, text), labels=False, refs=False)
    kind, d = same_as_host(pv, _write(tmp, "unsorted_dups", EDGE["unsorted_dups"]))
    assert d["ids"].tolist() == [3, 5, 7] and d["labels"] == b"GCAAC"          # first S line of id 7 wins
    assert d["ep1"].tolist() == [1, 4] and d["ep2"].tolist() == [3, 0] and d["steps"].tolist() == [0, 3, 4] and d["hap_len"].tolist() == [5]
    # a long segment (SWAR search across many words) and a walk much longer than one 16-byte chunk
    rng = random.Random(5)
    seq = "".join(rng.choice("ACGT") for _ in range(100_003))
    walk = ",".join(f"{rng.randint(1, 3)}{rng.choice('+-')}" for _ in range(5000))
    text = f"S\t1\t{seq}\nS\t2\tA\nS\t3\t{seq[:777]}\nL\t1\t+\t2\t+\t0M\nP\tlong#1#c\t{walk}\t*\nP\tshort\t2+\t*\n"
    kind, d = same_as_host(pv, _write(tmp, "long", text))
    assert kind == "ok" and d["step_off"].tolist() == [0, 5000, 5001]
    # with paths excluded a malformed P line is not looked at (from_gfa.cpp:245: only when inc_refs)
    kind, _ = same_as_host(pv, _write(tmp, "p_ignored", ERRORS["path_bad_orientation"]), labels=True, refs=False)
    assert kind == "ok"


def check_errors(pv, tmp):
    for name, text in ERRORS.items():
        kind, msg = same_as_host(pv, _write(tmp, name, text))
        assert kind == "err", (name, msg)
        assert msg.startswith("Invalid GFA '"), (name, msg)
    kind, msg = same_as_host(pv, _write(tmp, "unsupported_type", ERRORS["unsupported_type"]))
    assert msg.endswith("unsupported record type 'X' on line 4"), msg
    kind, msg = same_as_host(pv, _write(tmp, "s_one_tab", ERRORS["s_one_tab"]))
    assert msg.endswith("S record on line 3 is missing a sequence"), msg
    kind, msg = same_as_host(pv, os.path.join(tmp, "does_not_exist.gfa"))
    assert kind == "err" and msg.endswith("could not open file"), msg
    kind, msg = same_as_host(pv, tmp)          # a directory
    assert kind == "err" and msg.endswith("could not open file"), msg


def check_fuzz(pv, tmp, n_cases=150, seed=11):
    """Small random graphs written as GFA, then damaged at random bytes: the device parser and the host loader must agree on
    every outcome (identical arrays, or the identical error message)."""
    rng = random.Random(seed)
    n_err = 0
    for case in range(n_cases):
        g = synth.nested_bubbles(1000 + case, n_top=rng.randint(1, 4), max_depth=2, n_paths=rng.randint(1, 4))
        p = os.path.join(tmp, f"fuzz{case}.gfa")
        synth.write_gfa(g, p)
        text = bytearray(open(p, "rb").read())
        for _ in range(rng.choice([0, 1, 1, 2, 4])):
            i = rng.randrange(len(text))
            op = rng.randrange(4)
            if op == 0:
                text[i] = rng.choice(b"\t\n,+-0M*SLPX9 ")
            elif op == 1:
                del text[i]
            elif op == 2:
                text.insert(i, rng.choice(b"\t\n,+-7\r"))
            else:
                j = text.find(b"\n", i)
                if j > 0:
                    del text[i:j]
        with open(p, "wb") as f:
            f.write(text)
        kind, _ = same_as_host(pv, p)
        n_err += kind == "err"
    assert 0 < n_err < n_cases          # both outcomes were exercised


def check_large_synthetic(pv, tmp, n_bubbles=20_000):
    g = synth.snp_indel_chain(n_bubbles, 16, 42)
    p = os.path.join(tmp, "chain.gfa")
    synth.write_gfa_fast(g, p)
    kind, d = same_as_host(pv, p)
    assert kind == "ok"
    ep1, ep2 = api.edge_ports(g)
    assert np.array_equal(d["ep1"], ep1) and np.array_equal(d["ep2"], ep2) and np.array_equal(d["ids"], g["ids"])
    return d
