# SPDX-License-Identifier: MIT
"""Graph builders for the benchmark / parity configurations and the on-disk
formats either side of the path.

  * ``hmm3``              3-state left-to-right HMM of test/test_algorithms.jl:13-26 / demo.ipynb
  * ``chain``             a -> b -> c -> d chain of test/test_algorithms.jl:262-283
  * ``phone_loop``        cfg 1: fully connected loop of 3-state phone HMMs
  * ``numerator``         cfg 2: LinearFSM-shaped utterance graphs (examples/prepare-lfmmi-graphs.jl:25-65)
  * ``denominator``       cfg 3: Kaldi-chain 2-pdf topology x phonotactic LM, modelled on
                          misc/benchmark/den_fsm_wsj.txt
  * ``load_openfst_text`` the text format of misc/benchmark/*.txt (writer generatefsm.jl:42-57)

Every builder returns ``(fsm, pdfids)`` with 0-based pdf ids of the real states; weights are
natural-log probabilities in the payload dtype of ``K``.
"""
import numpy as np

from .fsm import FSM, renorm


def hmm3(K, nstates=3):
    """Left-to-right HMM: init on state 1, self-loop + forward arc, final on the last state,
    then ``renorm`` (test/test_algorithms.jl:13-26; demo.ipynb cell 13)."""
    S = nstates
    src = list(range(S)) + list(range(S - 1))
    dst = list(range(S)) + list(range(1, S))
    w = [K.one] * len(src)
    fsm = FSM.from_arrays(K, S, src, dst, w, [0], [K.one], [S - 1], [K.one], list(range(1, S + 1)))
    return renorm(fsm), np.arange(S)


def chain(K, nstates=4):
    """a -> b -> c -> d without self-loops (test/test_algorithms.jl:262-283)."""
    S = nstates
    fsm = FSM.from_arrays(K, S, list(range(S - 1)), list(range(1, S)), [K.one] * (S - 1), [0], [K.one],
                          [S - 1], [K.one], list(range(1, S + 1)))
    return fsm, np.arange(S)


def phone_loop(K, n_phones=33, states_per_phone=3):
    """cfg 1: P phones x 3-state L-R HMM (self-loop and forward arc, p = 1/2 each), the last
    state of every phone connects uniformly to the first state of every phone and to the final
    state; uniform initial distribution over phone entries; renormalised."""
    P, Q = n_phones, states_per_phone
    S = P * Q
    src, dst = [], []
    for p in range(P):
        for q in range(Q):
            s = p * Q + q
            src.append(s); dst.append(s)
            if q + 1 < Q:
                src.append(s); dst.append(s + 1)
            else:
                for p2 in range(P):
                    src.append(s); dst.append(p2 * Q)
    w = [K.one] * len(src)
    entries = [p * Q for p in range(P)]
    exits = [p * Q + Q - 1 for p in range(P)]
    fsm = FSM.from_arrays(K, S, src, dst, w, entries, [K.one] * P, exits, [K.one] * P, list(range(1, S + 1)))
    return renorm(fsm), np.arange(S)


def numerator(K, rng, n_pdf, n_phones=None, silprob=0.2, n_ctx=None):
    """cfg 2: one utterance's numerator graph.  A linear phone sequence where every phone is the
    2-state chain HMM of misc/benchmark/num_fsm_wsj.txt (entry state A, pdf 2p, no self-loop;
    looped state B, pdf 2p+1; A -> B and B -> B with 1/2, and A and B leave to the same
    successors with the other 1/2), with optional silence between word-like groups (probability
    ``silprob``) and 1-2 pronunciation variants (parallel branches) per group.  Rows sum to 1."""
    if n_phones is None:
        n_phones = int(rng.integers(20, 61))
    n_ctx = n_ctx or n_pdf // 2
    src, dst, w, pdfids = [], [], [], []
    ln = np.log

    def add_phone(ctx):
        a = len(pdfids); pdfids.append(2 * ctx)
        b = len(pdfids); pdfids.append(2 * ctx + 1)
        src.extend([a, b]); dst.extend([b, b]); w.extend([ln(0.5), ln(0.5)])
        return a, b

    def link(exits, targets):  # every exit state spends its remaining 1/2 on the targets
        for x in exits:
            for t, p in targets:
                src.append(x); dst.append(t); w.append(ln(0.5 * p))

    exits, init, k = None, [], 0
    while k < n_phones:
        group = int(min(rng.integers(1, 4), n_phones - k))
        variants = 1 + int(rng.random() < 0.3)
        starts, ends = [], []
        for _ in range(variants):
            prev = None
            for _ in range(group):
                a, b = add_phone(int(rng.integers(1, n_ctx)))
                if prev is None:
                    starts.append(a)
                else:
                    link(prev, [(a, 1.0)])
                prev = (a, b)
            ends.extend(prev)
        entry = [(s, 1.0 / len(starts)) for s in starts]
        if exits is None:
            init = entry
        elif silprob > 0:
            sa, sb = add_phone(0)  # silence phone
            link(exits, [(s, p * (1 - silprob)) for s, p in entry] + [(sa, silprob)])
            link((sa, sb), entry)
        else:
            link(exits, entry)
        exits = ends
        k += group
    S = len(pdfids)
    fsm = FSM.from_arrays(K, S, src, dst, np.asarray(w, K.dtype), [s for s, _ in init],
                          np.asarray([ln(p) for _, p in init], K.dtype), list(exits),
                          np.full(len(exits), ln(0.5), K.dtype), list(range(1, S + 1)))
    return fsm, np.asarray(pdfids)


def denominator(K, n_tokens=15000, n_pdf=3000, seed=303, mean_fanout=13.0, sigma=0.8, max_fanout=40):
    """cfg 3: synthetic Kaldi-chain-topology denominator graph.

    ``n_tokens`` tokens; token k owns state A = 2k (pdf 2p, no self-loop) and B = 2k+1 (pdf 2p+1,
    self-loop 1/2); A -> B with 1/2; A and B share one successor set of
    ``clip(round(LogNormal(ln mean_fanout, sigma)), 1, max_fanout)`` random tokens' A states with
    Dirichlet(1) LM probabilities scaled by 1/2 minus the final mass; the phone-in-context p of a
    token is Zipf(1.0) over n_pdf/2; ~31 % of tokens are final (both A and B); 2.5 % of A states
    are initial.  Rows sum to 1."""
    rng = np.random.default_rng(seed)
    n_ctx = n_pdf // 2
    zipf = 1.0 / np.arange(1, n_ctx + 1)
    zipf /= zipf.sum()
    ctx = rng.choice(n_ctx, size=n_tokens, p=zipf)
    fan = np.clip(np.rint(rng.lognormal(np.log(mean_fanout), sigma, n_tokens)), 1, min(max_fanout, n_tokens)).astype(np.int64)
    is_final = rng.random(n_tokens) < 0.31
    final_mass = np.where(is_final, rng.uniform(0.01, 0.1, n_tokens), 0.0)
    src, dst, w = [], [], []
    fin_idx, fin_w = [], []
    for k in range(n_tokens):
        succ = rng.choice(n_tokens, size=fan[k], replace=False)
        succ.sort()
        probs = rng.dirichlet(np.ones(fan[k])) * (0.5 - final_mass[k])
        a, b = 2 * k, 2 * k + 1
        for s in (a, b):
            src.append(np.full(fan[k] + 1, s)); dst.append(np.concatenate([[b], 2 * succ]))
            w.append(np.concatenate([[0.5], probs]))
        if is_final[k]:
            fin_idx.extend([a, b]); fin_w.extend([final_mass[k]] * 2)
    n_init = max(1, int(round(0.025 * n_tokens)))
    init_tok = np.sort(rng.choice(n_tokens, size=n_init, replace=False))
    init_p = rng.dirichlet(np.ones(n_init))
    pdfids = np.empty(2 * n_tokens, np.int64)
    pdfids[0::2] = 2 * ctx
    pdfids[1::2] = 2 * ctx + 1
    S = 2 * n_tokens
    with np.errstate(divide="ignore"):
        fsm = FSM.from_arrays(K, S, np.concatenate(src), np.concatenate(dst),
                              np.log(np.concatenate(w)).astype(K.dtype), 2 * init_tok,
                              np.log(init_p).astype(K.dtype), fin_idx, np.log(np.asarray(fin_w)).astype(K.dtype),
                              list(range(1, S + 1)))
    return fsm, pdfids


def load_openfst_text(path, K):
    """OpenFst text graph as written by misc/benchmark/generatefsm.jl:42-57: arcs
    ``src dst ilabel olabel cost`` (cost = -log w), initial weights as arcs from pseudo-state 0,
    final lines ``state cost``; states and pdf ids 1-based, the pdf of a state is the ilabel of
    its incoming arcs."""
    src, dst, w, ii, iw, fi, fw = [], [], [], [], [], [], []
    pdf = {}
    with open(path) as f:
        for line in f:
            t = line.split()
            if len(t) == 5:
                s, d, il, _, c = int(t[0]), int(t[1]), int(t[2]), int(t[3]), float(t[4])
                if pdf.setdefault(d, il) != il:
                    raise ValueError(f"state {d} has two different pdfs")
                if s == 0:
                    ii.append(d - 1); iw.append(-c)
                else:
                    src.append(s - 1); dst.append(d - 1); w.append(-c)
            elif len(t) == 2:
                fi.append(int(t[0]) - 1); fw.append(-float(t[1]))
            elif t:
                raise ValueError(f"unparsable line: {line!r}")
    S = max(max(src), max(dst), max(ii), max(fi)) + 1
    pdfids = np.array([pdf[s + 1] - 1 for s in range(S)], np.int64)
    fsm = FSM.from_arrays(K, S, src, dst, np.asarray(w, K.dtype), ii, np.asarray(iw, K.dtype), fi,
                          np.asarray(fw, K.dtype), list(range(1, S + 1)))
    return fsm, pdfids
