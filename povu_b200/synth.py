"""Deterministic synthetic pangenome graphs as flat arrays + GFA text.

A graph is a dict of numpy arrays:
  ids   u32[V]  segment ids, ascending
  seqs  list[str] or None
  v1,v2 u32[E]  vertex *indices* of each L line, in file order
  s1,s2 u8[E]   side of each endpoint (0 = left, 1 = right); `L a + b +` is (a,1)-(b,0)
  paths list of (name, steps u32[], strands u8[])   steps are vertex indices, strand 0 = '+'
Nothing here touches the GPU; generators are numpy (vectorised for the large configs).
"""
from __future__ import annotations

import numpy as np

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _rng(seed):
    return np.random.default_rng(seed)


def link(a, oa, b, ob):
    """GFA `L a oa b ob` -> (v1, side1, v2, side2).  '+' leaves through the right side and enters
    through the left (liteseq mapping)."""
    return (a, 1 if oa == "+" else 0, b, 0 if ob == "+" else 1)


def make_graph(ids, seqs, links, paths):
    links = np.asarray(links, dtype=np.int64).reshape(-1, 4)
    return {
        "ids": np.asarray(ids, dtype=np.uint32),
        "seqs": seqs,
        "v1": links[:, 0].astype(np.uint32),
        "s1": links[:, 1].astype(np.uint8),
        "v2": links[:, 2].astype(np.uint32),
        "s2": links[:, 3].astype(np.uint8),
        "paths": paths,
    }


def write_gfa(g, fp):
    ids = g["ids"]
    seqs = g["seqs"]
    out = ["H\tVN:Z:1.0\n"]
    for i, sid in enumerate(ids):
        out.append(f"S\t{sid}\t{seqs[i] if seqs is not None else 'A'}\n")
    o1 = np.where(g["s1"] == 1, "+", "-")
    o2 = np.where(g["s2"] == 0, "+", "-")
    a = ids[g["v1"]]
    b = ids[g["v2"]]
    for k in range(len(a)):
        out.append(f"L\t{a[k]}\t{o1[k]}\t{b[k]}\t{o2[k]}\t0M\n")
    for name, steps, strands in g["paths"]:
        st = ",".join(f"{ids[s]}{'+' if r == 0 else '-'}" for s, r in zip(steps, strands))
        out.append(f"P\t{name}\t{st}\t*\n")
    with open(fp, "w") as f:
        f.writelines(out)


def write_gfa_fast(g, fp):
    """Vectorised writer for the large configs (numpy char ops)."""
    ids = g["ids"].astype(np.int64)
    with open(fp, "w") as f:
        f.write("H\tVN:Z:1.0\n")
        seqs = g["seqs"]
        f.write("".join(f"S\t{i}\t{s}\n" for i, s in zip(ids.tolist(), seqs)))
        o1 = np.where(g["s1"] == 1, "+", "-").tolist()
        o2 = np.where(g["s2"] == 0, "+", "-").tolist()
        a = ids[g["v1"]].tolist()
        b = ids[g["v2"]].tolist()
        f.write("".join(f"L\t{x}\t{p}\t{y}\t{q}\t0M\n" for x, p, y, q in zip(a, o1, b, o2)))
        for name, steps, strands in g["paths"]:
            sid = ids[steps].tolist()
            sg = np.where(np.asarray(strands) == 0, "+", "-").tolist()
            f.write(f"P\t{name}\t" + ",".join(f"{x}{y}" for x, y in zip(sid, sg)) + "\t*\n")


# --------------------------------------------------------------------------------------
# cfg2: one chain of alternating SNP / indel bubbles (vectorised)
# --------------------------------------------------------------------------------------
def snp_indel_chain(n_bubbles=400_000, n_paths=16, seed=42, with_seqs=True, first_id=1):
    """SURVEY.md §8d cfg2 `snp_indel_1M`: backbone b_0..b_n, bubble i between b_i and b_{i+1}.
    Even i: SNP (two 1-base alleles); odd i: indel (one optional 1-8 bp segment).
    V = (n+1) + 2*ceil(n/2) + floor(n/2); E = 4*ceil(n/2) + 3*floor(n/2)."""
    rng = _rng(seed)
    n = n_bubbles
    n_snp = (n + 1) // 2
    n_ind = n // 2
    # vertex layout in id order along the chain: b_i, then alleles of bubble i
    per = np.where(np.arange(n) % 2 == 0, 3, 2)            # vertices contributed by block i (b_i + alleles)
    base = np.concatenate([[0], np.cumsum(per)])             # index of b_i ; b_n = base[n]
    V = int(base[n]) + 1
    ids = np.arange(first_id, first_id + V, dtype=np.uint32)
    b = base[:n]
    nxt = base[1:]
    snp = np.arange(n) % 2 == 0
    links = []
    # SNP: b->a0, b->a1, a0->nb, a1->nb ; indel: b->x, x->nb, b->nb
    bs, ns = b[snp], nxt[snp]
    bi, ni = b[~snp], nxt[~snp]
    one = np.ones_like
    L_snp = np.stack([
        np.stack([bs, one(bs), bs + 1, 0 * bs], 1),
        np.stack([bs, one(bs), bs + 2, 0 * bs], 1),
        np.stack([bs + 1, one(bs), ns, 0 * bs], 1),
        np.stack([bs + 2, one(bs), ns, 0 * bs], 1)], 1)      # [n_snp,4,4]
    L_ind = np.stack([
        np.stack([bi, one(bi), bi + 1, 0 * bi], 1),
        np.stack([bi + 1, one(bi), ni, 0 * bi], 1),
        np.stack([bi, one(bi), ni, 0 * bi], 1)], 1)          # [n_ind,3,4]
    # interleave in chain order so that L-line order follows the chain
    order_key = np.concatenate([np.repeat(np.arange(n)[snp], 4), np.repeat(np.arange(n)[~snp], 3)])
    allL = np.concatenate([L_snp.reshape(-1, 4), L_ind.reshape(-1, 4)])
    allL = allL[np.argsort(order_key, kind="stable")]
    seqs = None
    if with_seqs:
        lens = np.ones(V, dtype=np.int64)
        lens[base] = rng.integers(1, 17, size=n + 1)
        lens[b[~snp] + 1] = rng.integers(1, 9, size=n_ind)
        tot = int(lens.sum())
        letters = BASES[rng.integers(0, 4, size=tot)]
        # SNP alleles must differ: force allele1 = (allele0 + 1 + r) % 4
        off = np.concatenate([[0], np.cumsum(lens)])
        a0 = off[b[snp] + 1]
        a1 = off[b[snp] + 2]
        c0 = rng.integers(0, 4, size=n_snp)
        c1 = (c0 + 1 + rng.integers(0, 3, size=n_snp)) % 4
        letters[a0] = BASES[c0]
        letters[a1] = BASES[c1]
        s = letters.tobytes().decode()
        seqs = [s[off[i]:off[i + 1]] for i in range(V)]
    paths = []
    q = rng.uniform(0.05, 0.5, size=n)
    for p in range(n_paths):
        alt = (rng.uniform(size=n) < q) if p > 0 else np.zeros(n, dtype=bool)
        # per block: b_i, then chosen allele (SNP: a0/a1; indel: x unless alt -> skip)
        cnt = np.where(snp, 2, np.where(alt, 1, 2))
        offs = np.concatenate([[0], np.cumsum(cnt)])
        steps = np.empty(int(offs[-1]) + 1, dtype=np.uint32)
        steps[offs[:-1]] = b
        second = np.where(snp, b + 1 + alt.astype(np.int64), b + 1)
        has2 = cnt == 2
        steps[offs[:-1][has2] + 1] = second[has2]
        steps[-1] = base[n]
        sample = p // 2 + 1
        hap = p % 2 + 1
        paths.append((f"S{sample:03d}#{hap}#chr1", steps, np.zeros(len(steps), dtype=np.uint8)))
    g = make_graph(ids, seqs, allL, paths)
    return g


# --------------------------------------------------------------------------------------
# small random families for parity tests (python loops; sizes up to ~10^4)
# --------------------------------------------------------------------------------------
def nested_bubbles(seed, n_top=6, max_depth=3, n_paths=6, p_inv=0.15, p_del=0.2, p_multi=0.3,
                   shuffle_links=False, first_id=1):
    """Recursive bubble chains: every allele body may itself be a chain.  Optional inversion
    bubbles (`s+ y- t+` next to `s+ y+ t+`), deletion edges (s -> t directly) and multi-allelic
    sites.  Two tips.  Paths are random source-to-sink walks (PanSN names)."""
    rng = _rng(seed)
    seqs = []
    links = []

    def new_v():
        ln = int(rng.integers(1, 5))
        seqs.append("".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=ln)))
        return len(seqs) - 1

    # a "walk generator" returns list of (vertex, strand) choices: we record the structure as a tree
    def chain(depth, n_b):
        """returns (first_v, last_v, walker) ; walker(rng) -> list of (v, strand) from first to last inclusive"""
        nodes = [new_v()]
        bubbles = []
        for _ in range(n_b):
            s = nodes[-1]
            t = new_v()
            kind = rng.uniform()
            alleles = []
            n_all = 2 + (int(rng.integers(1, 3)) if rng.uniform() < p_multi else 0)
            if kind < p_inv:
                y = new_v()
                links.append(link(s, "+", y, "+")); links.append(link(y, "+", t, "+"))
                links.append(link(s, "+", y, "-")); links.append(link(y, "-", t, "+"))
                alleles = [("inv", y)]
            else:
                has_del = rng.uniform() < p_del
                for a in range(n_all - (1 if has_del else 0)):
                    if depth < max_depth and rng.uniform() < 0.4:
                        f, l, w = chain(depth + 1, int(rng.integers(1, 4)))
                        links.append(link(s, "+", f, "+")); links.append(link(l, "+", t, "+"))
                        alleles.append(("sub", w))
                    else:
                        x = new_v()
                        links.append(link(s, "+", x, "+")); links.append(link(x, "+", t, "+"))
                        alleles.append(("seg", x))
                if has_del:
                    links.append(link(s, "+", t, "+"))
                    alleles.append(("del", None))
            bubbles.append(alleles)
            nodes.append(t)

        def walker(r):
            out = [(nodes[0], 0)]
            for i, alleles in enumerate(bubbles):
                k = alleles[int(r.integers(0, len(alleles)))]
                if k[0] == "inv":
                    out.append((k[1], int(r.uniform() < 0.3)))
                elif k[0] == "sub":
                    out.extend(k[1](r))
                elif k[0] == "seg":
                    out.append((k[1], 0))
                out.append((nodes[i + 1], 0))
            return out
        return nodes[0], nodes[-1], walker

    _, _, w = chain(0, n_top)
    V = len(seqs)
    paths = []
    for p in range(n_paths):
        walk = w(rng)
        steps = np.array([x for x, _ in walk], dtype=np.uint32)
        strands = np.array([s for _, s in walk], dtype=np.uint8)
        paths.append((f"S{p // 2 + 1:03d}#{p % 2 + 1}#chr1", steps, strands))
    links = np.asarray(links, dtype=np.int64).reshape(-1, 4)
    if shuffle_links:
        links = links[rng.permutation(len(links))]
    ids = np.arange(first_id, first_id + V, dtype=np.uint32)
    return make_graph(ids, seqs, links, paths)


def wild(seed, V=12, E=18, p_loop=0.05, n_comp=1):
    """Arbitrary bidirected multigraphs: random sides, self-loops, parallel edges, several
    components, components with and without tips.  No paths."""
    rng = _rng(seed)
    links = []
    per = max(1, V // n_comp)
    for _ in range(E):
        c = int(rng.integers(0, n_comp))
        lo, hi = c * per, min(V, (c + 1) * per) if c < n_comp - 1 else V
        a = int(rng.integers(lo, hi))
        b = a if rng.uniform() < p_loop else int(rng.integers(lo, hi))
        links.append((a, int(rng.integers(0, 2)), b, int(rng.integers(0, 2))))
    ids = np.arange(1, V + 1, dtype=np.uint32)
    return make_graph(ids, ["A"] * V, links, [])


def concat_components(graphs):
    """Disjoint union (ids renumbered consecutively) -> multi-component graph."""
    ids, seqs, links, paths = [], [], [], []
    off = 0
    for ci, g in enumerate(graphs):
        V = len(g["ids"])
        seqs.extend(g["seqs"] if g["seqs"] is not None else ["A"] * V)
        L = np.stack([g["v1"].astype(np.int64) + off, g["s1"], g["v2"].astype(np.int64) + off, g["s2"]], 1)
        links.append(L)
        for name, st, sr in g["paths"]:
            nm = name.rsplit("#", 1)[0] + f"#chr{ci + 1}" if "#" in name else f"{name}_{ci + 1}"
            paths.append((nm, st + off, sr))
        off += V
    ids = np.arange(1, off + 1, dtype=np.uint32)
    return make_graph(ids, seqs, np.concatenate(links) if links else np.zeros((0, 4)), paths)
