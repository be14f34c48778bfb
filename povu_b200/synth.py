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
    g["inst_base"] = base.astype(np.int64)      # backbone vertex of every bubble (+ the last one): where the chain can be cut
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


# --------------------------------------------------------------------------------------
# cfg3 / cfg4 / cfg5: large graphs by tiling small templates (vectorised)
# --------------------------------------------------------------------------------------
def tower(depth, seed=0, n_walks=8):
    """PVST depth stress: `depth` bubbles nested inside each other.  Level d: s_d -> x_d -> t_d (alt allele) and
    s_d -> s_{d+1} ... t_{d+1} -> t_d (the nested allele); innermost s_D -> y -> t_D.  Walk 0 goes all the way in;
    the others leave at a random level.  Returns a template dict (see `tiled`)."""
    rng = _rng(seed)
    D = depth
    # vertex layout: s_0..s_D, y, t_D..t_0, x_0..x_{D-1}
    s = np.arange(0, D + 1); y = D + 1; t = (2 * D + 2) - np.arange(0, D + 1); x = 2 * D + 3 + np.arange(0, D)
    V = 3 * D + 3
    links = []
    for d in range(D):
        links += [link(int(s[d]), "+", int(x[d]), "+"), link(int(x[d]), "+", int(t[d]), "+"),
                  link(int(s[d]), "+", int(s[d + 1]), "+"), link(int(t[d + 1]), "+", int(t[d]), "+")]
    links += [link(int(s[D]), "+", y, "+"), link(y, "+", int(t[D]), "+")]
    seqs = ["".join("ACGT"[int(c)] for c in rng.integers(0, 4, size=int(rng.integers(1, 4)))) for _ in range(V)]
    walks = []
    for w in range(n_walks):
        k = D if w == 0 else int(rng.integers(0, D + 1))
        mid = [y] if k == D else [int(x[k])]
        st = list(s[:k + 1]) + mid + list(t[:k + 1][::-1])
        walks.append((np.array(st, dtype=np.uint32), np.zeros(len(st), dtype=np.uint8)))
    return {"V": V, "seqs": seqs, "links": np.asarray(links, dtype=np.int64).reshape(-1, 4), "walks": walks}


def template_from(g):
    """A two-tip graph with source-to-sink paths -> tiling template (its paths become the walk choices)."""
    L = np.stack([g["v1"].astype(np.int64), g["s1"].astype(np.int64), g["v2"].astype(np.int64), g["s2"].astype(np.int64)], 1)
    return {"V": len(g["ids"]), "seqs": g["seqs"], "links": L, "walks": [(np.asarray(p[1], np.uint32), np.asarray(p[2], np.uint8)) for p in g["paths"]]}


def tiled(templates, n_instances, n_paths, seed=42, n_components=1, comp_weights=None, first_id=1, ref_walk0=True, tpl_weights=None):
    """Chains `n_instances` randomly chosen template instances (sink of one -> source of the next, `+ +`) into
    `n_components` components (sizes proportional to comp_weights).  Path p of component c picks one of the
    template's walks per instance (path 0 always walk 0 when ref_walk0).  Paths are named S%03d#{1,2}#chr<c>."""
    rng = _rng(seed)
    K = len(templates)
    if tpl_weights is None:
        tpl = rng.integers(0, K, size=n_instances)
    else:                                                        # templates drawn with the given relative frequencies
        w_ = np.asarray(tpl_weights, dtype=np.float64)
        tpl = np.minimum(np.searchsorted(np.cumsum(w_) / w_.sum(), rng.uniform(size=n_instances), side="right"), K - 1)
    tV = np.array([t["V"] for t in templates], dtype=np.int64)
    base = np.concatenate([[0], np.cumsum(tV[tpl])])            # first vertex of every instance
    V = int(base[-1])
    # components: contiguous runs of instances
    w = np.asarray(comp_weights if comp_weights is not None else np.ones(n_components), dtype=np.float64)
    cuts = np.round(np.cumsum(w) / w.sum() * n_instances).astype(np.int64)
    cuts[-1] = n_instances
    comp_of = np.searchsorted(cuts, np.arange(n_instances), side="right")
    # vertices / labels
    seq_pool = [np.array(t["seqs"], dtype=object) for t in templates]
    seqs = np.empty(V, dtype=object)
    links = []
    src = np.array([int(t["walks"][0][0][0]) for t in templates]); snk = np.array([int(t["walks"][0][0][-1]) for t in templates])
    for k in range(K):
        inst = np.nonzero(tpl == k)[0]
        if len(inst) == 0:
            continue
        b = base[inst]
        idx = (b[:, None] + np.arange(tV[k])[None, :]).ravel()
        seqs[idx] = np.tile(seq_pool[k], len(inst))
        L = templates[k]["links"]
        LL = np.repeat(L[None, :, :], len(inst), axis=0).copy()
        LL[:, :, 0] += b[:, None]; LL[:, :, 2] += b[:, None]
        links.append((inst.repeat(len(L)), LL.reshape(-1, 4)))
    # joins between consecutive instances of the same component
    i = np.arange(n_instances - 1)
    same = comp_of[i] == comp_of[i + 1]
    j = i[same]
    J = np.stack([base[j] + snk[tpl[j]], np.ones(len(j), np.int64), base[j + 1] + src[tpl[j + 1]], np.zeros(len(j), np.int64)], 1)
    key = np.concatenate([k_ for k_, _ in links] + [j + 0.5])
    allL = np.concatenate([l_ for _, l_ in links] + [J])
    allL = allL[np.argsort(key, kind="stable")]                  # L lines follow the chain
    # paths
    nW = np.array([len(t["walks"]) for t in templates])
    wl = [[len(wk[0]) for wk in t["walks"]] for t in templates]
    paths = []
    for c in range(n_components):
        inst = np.nonzero(comp_of == c)[0]
        if len(inst) == 0:
            continue
        for p in range(n_paths):
            choice = rng.integers(0, 1 << 30, size=len(inst)) % nW[tpl[inst]]
            if p == 0 and ref_walk0:
                choice[:] = 0
            lens = np.array([wl[k][w_] for k, w_ in zip(tpl[inst].tolist(), choice.tolist())], dtype=np.int64) if len(inst) < 2000 else None
            if lens is None:
                lens = np.zeros(len(inst), dtype=np.int64)
                for k in range(K):
                    for w_ in range(nW[k]):
                        m = (tpl[inst] == k) & (choice == w_)
                        lens[m] = wl[k][w_]
            off = np.concatenate([[0], np.cumsum(lens)])
            steps = np.empty(int(off[-1]), dtype=np.uint32); strands = np.empty(int(off[-1]), dtype=np.uint8)
            for k in range(K):
                for w_ in range(nW[k]):
                    m = np.nonzero((tpl[inst] == k) & (choice == w_))[0]
                    if len(m) == 0:
                        continue
                    ws, wr = templates[k]["walks"][w_]
                    pos = (off[m][:, None] + np.arange(len(ws))[None, :]).ravel()
                    steps[pos] = (base[inst[m]][:, None] + ws[None, :].astype(np.int64)).ravel().astype(np.uint32)
                    strands[pos] = np.tile(wr, len(m))
            paths.append((f"S{p // 2 + 1:03d}#{p % 2 + 1}#chr{c + 1}", steps, strands))
    ids = np.arange(first_id, first_id + V, dtype=np.uint32)
    g = make_graph(ids, seqs.tolist(), allL, paths)
    g["inst_base"] = base.astype(np.int64)      # first vertex of every template instance (+ V): where the chain can be cut
    return g


def tiled_flat(templates, n_instances, n_paths, seed=42, chrom="chr1", first_id=1, tpl_weights=None):
    """One chain-shaped component like `tiled`, but built as FLAT numpy arrays only (no Python object per segment or step): the
    form a chromosome-scale component needs (10^7 segments x 90 walks).  Returns ids, seg_len, lab_off / lab (concatenated
    sequences), v1 / s1 / v2 / s2, names, step_off and packed steps ((vertex << 1) | strand), inst_base.  Path 0 takes walk 0
    of every instance.  Not the same random stream as `tiled` (its own graphs: `window_flat` cuts the reference's share)."""
    rng = _rng(seed)
    K = len(templates)
    if tpl_weights is None:
        tpl = rng.integers(0, K, size=n_instances)
    else:
        w_ = np.asarray(tpl_weights, dtype=np.float64)
        tpl = np.minimum(np.searchsorted(np.cumsum(w_) / w_.sum(), rng.uniform(size=n_instances), side="right"), K - 1)
    tV = np.array([t["V"] for t in templates], dtype=np.int64)
    base = np.concatenate([[0], np.cumsum(tV[tpl])])
    V = int(base[-1])
    # per-template pools: segment lengths / sequences, links, walks -- all flat
    t_seglen = [np.array([len(x) for x in t["seqs"]], dtype=np.int64) for t in templates]
    t_lab = [np.frombuffer("".join(t["seqs"]).encode(), dtype=np.uint8) for t in templates]
    t_labn = np.array([len(x) for x in t_lab], dtype=np.int64)
    seg_len = np.empty(V, dtype=np.uint32)
    lab_base = np.concatenate([[0], np.cumsum(t_labn[tpl])])      # first base of every instance
    lab = np.empty(int(lab_base[-1]), dtype=np.uint8)
    links = []
    for k in range(K):
        inst = np.nonzero(tpl == k)[0]
        if len(inst) == 0:
            continue
        b = base[inst]
        seg_len[(b[:, None] + np.arange(tV[k])[None, :]).ravel()] = np.tile(t_seglen[k].astype(np.uint32), len(inst))
        lab[(lab_base[inst][:, None] + np.arange(t_labn[k])[None, :]).ravel()] = np.tile(t_lab[k], len(inst))
        L = templates[k]["links"]
        LL = np.repeat(L[None, :, :], len(inst), axis=0).copy()
        LL[:, :, 0] += b[:, None]; LL[:, :, 2] += b[:, None]
        links.append((inst.repeat(len(L)), LL.reshape(-1, 4)))
    src = np.array([int(t["walks"][0][0][0]) for t in templates]); snk = np.array([int(t["walks"][0][0][-1]) for t in templates])
    j = np.arange(n_instances - 1)
    J = np.stack([base[j] + snk[tpl[j]], np.ones(len(j), np.int64), base[j + 1] + src[tpl[j + 1]], np.zeros(len(j), np.int64)], 1)
    key = np.concatenate([k_ for k_, _ in links] + [j + 0.5])
    allL = np.concatenate([l_ for _, l_ in links] + [J])
    allL = allL[np.argsort(key, kind="stable")]
    del links, key
    lab_off = np.concatenate([[0], np.cumsum(seg_len, dtype=np.uint64)]).astype(np.uint64)
    # walks: one flat pool, walk w of template k = pool[wstart[woff[k] + w] : + wlen[..]] (packed (vertex << 1) | strand, template-local)
    nW = np.array([len(t["walks"]) for t in templates], dtype=np.int64)
    woff = np.concatenate([[0], np.cumsum(nW)])
    wlen = np.array([len(wk[0]) for t in templates for wk in t["walks"]], dtype=np.int64)
    wstart = np.concatenate([[0], np.cumsum(wlen)])[:-1]
    pool = np.concatenate([(wk[0].astype(np.int64) << 1) | wk[1].astype(np.int64) for t in templates for wk in t["walks"]])
    names, parts, step_off = [], [], [0]
    for p in range(n_paths):
        choice = rng.integers(0, 1 << 30, size=n_instances) % nW[tpl]
        if p == 0:
            choice[:] = 0
        w = woff[tpl] + choice
        lens = wlen[w]
        off = np.concatenate([[0], np.cumsum(lens)])
        tot = int(off[-1])
        within = np.arange(tot, dtype=np.int64) - np.repeat(off[:-1], lens)
        st = pool[np.repeat(wstart[w], lens) + within] + (np.repeat(base[:-1], lens) << 1)
        parts.append(st.astype(np.uint32)); step_off.append(step_off[-1] + tot)
        names.append(f"S{p // 2 + 1:03d}#{p % 2 + 1}#{chrom}")
    return {"ids": np.arange(first_id, first_id + V, dtype=np.uint32), "seg_len": seg_len, "lab_off": lab_off, "lab": lab,
            "v1": allL[:, 0].astype(np.uint32), "s1": allL[:, 1].astype(np.uint8), "v2": allL[:, 2].astype(np.uint32), "s2": allL[:, 3].astype(np.uint8),
            "names": names, "step_off": np.asarray(step_off, dtype=np.uint64), "steps": np.concatenate(parts) if parts else np.empty(0, np.uint32), "inst_base": base.astype(np.int64)}


def window_flat(f, lo, hi):
    """vertices [lo, hi) of a `tiled_flat` component as an ordinary graph dict (strings and per-path arrays: small), with the
    original ids -- what `window` does for `tiled` graphs"""
    keep = (f["v1"] >= lo) & (f["v1"] < hi) & (f["v2"] >= lo) & (f["v2"] < hi)
    L = np.stack([f["v1"][keep].astype(np.int64) - lo, f["s1"][keep], f["v2"][keep].astype(np.int64) - lo, f["s2"][keep]], 1)
    lab = f["lab"].tobytes() if len(f["lab"]) < (1 << 26) else None
    seqs = []
    for v in range(lo, hi):
        a, b = int(f["lab_off"][v]), int(f["lab_off"][v + 1])
        seqs.append((lab[a:b] if lab is not None else f["lab"][a:b].tobytes()).decode())
    paths = []
    for k, name in enumerate(f["names"]):
        st = f["steps"][int(f["step_off"][k]):int(f["step_off"][k + 1])]
        vtx = st >> 1
        a, b = int(np.searchsorted(vtx, lo, side="left")), int(np.searchsorted(vtx, hi, side="left"))      # chain-shaped: steps ascend with the instances
        m = np.nonzero((vtx >= lo) & (vtx < hi))[0] if not (np.all(vtx[a:b] >= lo) and np.all(vtx[a:b] < hi) and (a == 0 or vtx[a - 1] < lo)) else None
        if m is not None:
            a, b = (int(m[0]), int(m[-1]) + 1) if len(m) else (0, 0)
            assert b - a == len(m), "the walk leaves and re-enters the window"
        paths.append((name, (vtx[a:b].astype(np.int64) - lo).astype(np.uint32), (st[a:b] & 1).astype(np.uint8)))
    return make_graph(f["ids"][lo:hi], seqs, L, paths)


def inversion_unit(seed=0, inner=1):
    """A block that paths may traverse on the reverse strand: u -> [s -> (x | y) -> t] -> v plus the two links of the
    inverted traversal (u+ -> t-, s- -> v+).  A path that walks t-, x-, s- crosses the region (s, t) as a whole-region
    inversion of the reference's s+, x+, t+: an SNE pin pair, a reverse-complement match, one SUBR record.  `inner` chains
    that many bubbles inside the block.  Walk 0 (the reference) is forward; 1 walk in 5 is inverted (p = 0.2)."""
    rng = _rng(1000 + seed)
    seqs, links = [], []

    def new_v():
        seqs.append("".join("ACGT"[int(c)] for c in rng.integers(0, 4, size=int(rng.integers(1, 5)))))
        return len(seqs) - 1

    u = new_v()
    chain = [new_v()]
    alleles = []
    for _ in range(inner):
        s_ = chain[-1]; x, y, t = new_v(), new_v(), new_v()
        links += [link(s_, "+", x, "+"), link(s_, "+", y, "+"), link(x, "+", t, "+"), link(y, "+", t, "+")]
        alleles.append((x, y)); chain.append(t)
    v = new_v()
    first, last = chain[0], chain[-1]
    links += [link(u, "+", first, "+"), link(last, "+", v, "+"), link(u, "+", last, "-"), link(first, "-", v, "+")]

    def walk(pick, inverted):
        mid = [chain[0]]
        for k, (x, y) in enumerate(alleles):
            mid += [x if pick[k] == 0 else y, chain[k + 1]]
        st = [(m, 0) for m in mid]
        if inverted:
            st = [(m, 1) for m, _ in reversed(st)]
        st = [(u, 0)] + st + [(v, 0)]
        return np.array([a for a, _ in st], dtype=np.uint32), np.array([b for _, b in st], dtype=np.uint8)

    zeros = [0] * inner
    walks = [walk(zeros, False), walk([1] * inner, False), walk([int(c) for c in rng.integers(0, 2, size=inner)], False),
             walk([int(c) for c in rng.integers(0, 2, size=inner)], False), walk(zeros if seed % 2 == 0 else [1] * inner, True)]
    return {"V": len(seqs), "seqs": seqs, "links": np.asarray(links, dtype=np.int64).reshape(-1, 4), "walks": walks}


def cfg3_templates(seed=7, n_nested=24, tower_depth=64, p_inv=0.0, n_repeat=0):
    """nested templates with depth mostly <= 3, some 4-8, plus one tower (inversion bubbles only
    when p_inv > 0 -- their SUBR records need the SNE pass; n_repeat VNTR-like repeat-unit templates give tangled regions)."""
    t = []
    for k in range(n_nested):
        depth = 2 + (k % 3) if k % 8 else 5 + (k % 4)
        t.append(template_from(nested_bubbles(seed * 1000 + k, n_top=4 + k % 5, max_depth=depth, n_paths=8, p_inv=p_inv)))
    for k in range(n_repeat):
        t.append(template_from(repeat_loops(2 * (seed * 100 + k), n_units=2, n_paths=8, max_laps=3)))       # even seed: walk 0 (the reference) takes one lap
    if tower_depth:
        t.append(tower(tower_depth, seed))
    return t


def window(g, lo, hi):
    """Vertices [lo, hi) of a chain-shaped graph with their ORIGINAL ids, the links among them and, of every path, its steps
    inside the window (a chain-shaped walk crosses the window once, so they are one run).  With lo / hi at instance
    boundaries of `tiled` (or backbone vertices of `snp_indel_chain`) every region strictly inside the window is the same
    region as in the whole graph: same ids, same alleles, same genotypes; POS is shifted by the bases of the reference walk
    before the window (`window_pos_shift`)."""
    keep = (g["v1"] >= lo) & (g["v1"] < hi) & (g["v2"] >= lo) & (g["v2"] < hi)
    L = np.stack([g["v1"][keep].astype(np.int64) - lo, g["s1"][keep], g["v2"][keep].astype(np.int64) - lo, g["s2"][keep]], 1)
    paths = []
    for name, st, sr in g["paths"]:
        m = np.nonzero((st >= lo) & (st < hi))[0]
        if len(m) == 0:
            paths.append((name, st[:0], sr[:0])); continue
        a, b = int(m[0]), int(m[-1]) + 1
        assert b - a == len(m), "the walk leaves and re-enters the window"
        paths.append((name, (st[a:b].astype(np.int64) - lo).astype(np.uint32), sr[a:b]))
    return make_graph(g["ids"][lo:hi], g["seqs"][lo:hi] if g["seqs"] is not None else None, L, paths)


def window_pos_shift(g, lo, hi, path=0):
    """bases of path `path`'s walk before its first step inside [lo, hi): POS(whole graph) = POS(window) + shift"""
    _, st, _ = g["paths"][path]
    m = np.nonzero((st >= lo) & (st < hi))[0]
    if len(m) == 0:
        return 0
    lens = np.fromiter((len(g["seqs"][v]) for v in st[:int(m[0])].tolist()), dtype=np.int64, count=int(m[0])) if g["seqs"] is not None else np.ones(int(m[0]), np.int64)
    return int(lens.sum())


def truncate(g, n_vertices):
    """Prefix of a chain-shaped graph: vertices [0, n), the links among them, every path cut at its first step
    outside.  With n at an instance boundary of `tiled` / a backbone vertex of `snp_indel_chain`, every region
    strictly inside the prefix is identical to the same region of the whole graph (same ids, same POS)."""
    keep = (g["v1"] < n_vertices) & (g["v2"] < n_vertices)
    L = np.stack([g["v1"][keep].astype(np.int64), g["s1"][keep], g["v2"][keep].astype(np.int64), g["s2"][keep]], 1)
    paths = []
    for name, st, sr in g["paths"]:
        out = np.nonzero(st >= n_vertices)[0]
        k = int(out[0]) if len(out) else len(st)
        paths.append((name, st[:k], sr[:k]))
    return make_graph(g["ids"][:n_vertices], g["seqs"][:n_vertices] if g["seqs"] is not None else None, L, paths)


def repeat_loops(seed, n_units=6, n_paths=6, max_laps=4, first_id=1):
    """Tangled regions: a backbone whose units `r1 -> (x | y | -) -> r2` carry a back link r2 -> r1, and paths that
    run through a unit 1..max_laps times picking an allele per lap (a VNTR-like repeat, as in LPA's KIV-2).
    Path 0 takes one lap of allele x everywhere unless `seed` is odd (then it loops too)."""
    rng = _rng(seed)
    seqs, links, units = [], [], []

    def new_v():
        seqs.append("".join("ACGT"[int(c)] for c in rng.integers(0, 4, size=int(rng.integers(1, 5)))))
        return len(seqs) - 1

    prev = new_v()
    for _ in range(n_units):
        r1, x, y, r2, nxt = new_v(), new_v(), new_v(), new_v(), new_v()
        links += [link(prev, "+", r1, "+"), link(r1, "+", x, "+"), link(r1, "+", y, "+"), link(x, "+", r2, "+"), link(y, "+", r2, "+"),
                  link(r2, "+", r1, "+"), link(r2, "+", nxt, "+")]
        has_del = rng.uniform() < 0.4
        if has_del:
            links.append(link(r1, "+", r2, "+"))
        units.append((prev, r1, x, y, r2, nxt, has_del))
        prev = nxt
    paths = []
    for p in range(n_paths):
        st = [units[0][0]]
        for (_, r1, x, y, r2, nxt, has_del) in units:
            laps = 1 if (p == 0 and seed % 2 == 0) else int(rng.integers(1, max_laps + 1))
            for _ in range(laps):
                st.append(r1)
                c = 0 if (p == 0 and seed % 2 == 0) else int(rng.integers(0, 3 if has_del else 2))
                if c < 2:
                    st.append(x if c == 0 else y)
                st.append(r2)
            st.append(nxt)
        paths.append((f"S{p // 2 + 1:03d}#{p % 2 + 1}#chr1", np.array(st, dtype=np.uint32), np.zeros(len(st), dtype=np.uint8)))
    return make_graph(np.arange(first_id, first_id + len(seqs), dtype=np.uint32), seqs, links, paths)


def cyclic_units(seed, n_units=5, n_paths=6, max_laps=3, first_id=1):
    """Regions whose BFS/Kahn vertex order fails (a cycle strictly inside a bubble), so the reference falls back to
    its haplotype-lap order (graph.cpp:194-395): units `s -> x1 -> .. -> xk -> t` (k = 2..3) with a back link
    xk -> x1, an alternative branch `s -> w -> t` and sometimes a chord `x1 -> t`; paths run the cycle 1..max_laps
    times or take the branch.  A leading backbone vertex keeps every lap away from step 0 (the reference reads
    v_ids[-1] there)."""
    rng = _rng(seed)
    seqs, links, units = [], [], []

    def new_v():
        seqs.append("".join("ACGT"[int(c)] for c in rng.integers(0, 4, size=int(rng.integers(1, 5)))))
        return len(seqs) - 1

    prev = new_v()
    for _ in range(n_units):
        k = int(rng.integers(2, 4))
        s_ = new_v(); xs = [new_v() for _ in range(k)]; w = new_v(); t = new_v(); nxt = new_v()
        links += [link(prev, "+", s_, "+"), link(s_, "+", xs[0], "+")]
        links += [link(xs[i], "+", xs[i + 1], "+") for i in range(k - 1)]
        links += [link(xs[-1], "+", xs[0], "+"), link(xs[-1], "+", t, "+"), link(s_, "+", w, "+"), link(w, "+", t, "+"), link(t, "+", nxt, "+")]
        chord = rng.uniform() < 0.3
        if chord:
            links.append(link(xs[0], "+", t, "+"))
        units.append((s_, xs, w, t, nxt, chord))
        prev = nxt
    paths = []
    for p in range(n_paths):
        st = [0]
        for (s_, xs, w, t, nxt, chord) in units:
            st.append(s_)
            c = 0 if p == 0 else int(rng.integers(0, 3))
            if c == 2:
                st.append(w)
            else:
                laps = 1 if (p == 0 and seed % 2 == 0) else int(rng.integers(1, max_laps + 1))
                for _ in range(laps):
                    st += xs
                if chord and c == 1:
                    st.append(xs[0])
            st += [t, nxt]
        paths.append((f"S{p // 2 + 1:03d}#{p % 2 + 1}#chr1", np.array(st, dtype=np.uint32), np.zeros(len(st), dtype=np.uint8)))
    return make_graph(np.arange(first_id, first_id + len(seqs), dtype=np.uint32), seqs, links, paths)


def reverse_paths(g, which):
    """The same graph with the paths in `which` walked backwards (steps reversed, strands flipped): a path on the
    reverse strand needs no extra links (`L a + b +` is also `L b - a -`).  Regions such a path traverses with the
    reference's alleles become whole-region inversions -> SNE pin pairs -> SUBR records."""
    paths = []
    for i, (name, st, sr) in enumerate(g["paths"]):
        if i in which:
            paths.append((name, np.ascontiguousarray(st[::-1]), np.ascontiguousarray(1 - sr[::-1]).astype(np.uint8)))
        else:
            paths.append((name, st, sr))
    out = dict(g)
    out["paths"] = paths
    return out


def permute_vertices(g, seed):
    """The same graph with its vertices renumbered at random (ids stay 1..V in index order): segment ids then no longer follow
    the backbone -- the decompose heuristics that look at id order must stay exact (they may only cost extra rounds)."""
    rng = _rng(seed)
    V = len(g["ids"])
    perm = rng.permutation(V).astype(np.int64)          # old index -> new index
    inv = np.empty(V, dtype=np.int64); inv[perm] = np.arange(V)
    seqs = [g["seqs"][int(i)] for i in inv] if g["seqs"] is not None else None
    L = np.stack([perm[g["v1"].astype(np.int64)], g["s1"].astype(np.int64), perm[g["v2"].astype(np.int64)], g["s2"].astype(np.int64)], 1)
    paths = [(n, perm[st.astype(np.int64)].astype(np.uint32), sr) for n, st, sr in g["paths"]]
    return make_graph(g["ids"], seqs, L, paths)
