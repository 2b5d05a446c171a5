"""Synthetic TREC-QA-shaped inputs for the MMS hot path (numpy, host side).

Shapes, distributions and seeds follow SURVEY.md 8(d), which in turn takes the
reference's own constants where they exist (file:line relative to the reference's
examples/trec_qa_w2v_mms/do_trec_qa_clean.py):

* seed 22 (:59); embedding table W ~ U(-0.08, 0.08) (:387), bias 0;
* sentences of length lq ~ U{3..20}, la ~ U{5..40}, token ids ~ U{0..V-3},
  centre-padded to L with the pad id V-1 exactly like ``vocab_transform_embed``
  (:184-203: pad_b = (L - l) // 2, truncate to L); indices are stored as
  float32 because the reference's HDF5 data layer feeds Embed float blobs
  (hdf5_data_layer.cpp:150-155, embed_layer.cpp:142);
* M ~ U(-0.1, 0.1) for kernel tests (the example's default filler is all-zero),
  B = 0; labels y ~ Bernoulli(0.15) (README.md:22-23: 248 / 1442 positive).

Named configurations C1..C5 are BASELINE.json's ``configs``.
"""
import numpy as np

SEED = 22

CONFIGS = {
    # name: dict(N, L, D, mc, V)
    "c1": dict(N=50, L=40, D=50, mc=4, V=60002),
    "c2": dict(N=50, L=40, D=300, mc=4, V=60002),
    "c3": dict(N=4096, L=40, D=300, mc=4, V=60002),
    # sentence-vector / reranking configurations (SimMatrix)
    "c4": dict(Nq=1000, Nc=1000000, K=1024),
    "c5": dict(N=16384, K1=1024, K2=1024, modalities=4),
}

C2_PAIRS_PER_EPOCH = 53417  # do_trec_qa_clean.py:37 (train-all, clean)


def pad_sentence(tokens, L, pad_id):
    """Centre-pad / truncate like vocab_transform_embed (do_trec_qa_clean.py:184-203)."""
    tokens = list(tokens)[:L]
    pad_b = (L - len(tokens)) // 2
    out = np.full(L, pad_id, dtype=np.int64)
    out[pad_b:pad_b + len(tokens)] = tokens
    return out


def make_indices(rng, N, L, V, len_lo, len_hi):
    """(N, L) float32 token ids, centre-padded with id V-1."""
    idx = np.full((N, L), V - 1, dtype=np.int64)
    lens = rng.integers(len_lo, len_hi + 1, size=N)
    for n in range(N):
        toks = rng.integers(0, V - 2, size=int(lens[n]))
        idx[n] = pad_sentence(toks, L, V - 1)
    return idx.astype(np.float32)


def make_qa_batch(N, L=40, D=300, mc=4, V=60002, seed=SEED, with_table=True, dtype=np.float32):
    """One synthetic batch for Embed x2 -> SimCross(mode 2).

    Returns a dict with idx_q, idx_a (N, L) float ids; W (V, D), b (D,);
    M (mc, D, D), B (mc, L, L); dS (N, mc, L, L) upstream gradient ~ U(-1,1)/count;
    y (N,) labels.
    """
    rng = np.random.default_rng(seed)
    out = {}
    out["idx_q"] = make_indices(rng, N, L, V, 3, 20)
    out["idx_a"] = make_indices(rng, N, L, V, 5, 40)
    if with_table:
        out["W"] = rng.uniform(-0.08, 0.08, size=(V, D)).astype(dtype)
        out["b"] = np.zeros(D, dtype=dtype)
    out["M"] = rng.uniform(-0.1, 0.1, size=(mc, D, D)).astype(dtype)
    out["B"] = np.zeros((mc, L, L), dtype=dtype)
    count = N * mc * L * L
    out["dS"] = (rng.uniform(-1.0, 1.0, size=(N, mc, L, L)) / count).astype(dtype)
    out["y"] = (rng.random(N) < 0.15).astype(dtype)
    return out


def make_sentence_vectors(N, K1=1024, K2=1024, seed=SEED, dtype=np.float32):
    """C5-style inputs for SimMatrix: q^, a^ ~ tanh(N(0,1)); W ~ xavier (fan_in)."""
    rng = np.random.default_rng(seed)
    q = np.tanh(rng.standard_normal((N, K1))).astype(dtype)
    a = np.tanh(rng.standard_normal((N, K2))).astype(dtype)
    scale = np.sqrt(3.0 / K1)  # caffe "xavier": U(-sqrt(3/fan_in), +sqrt(3/fan_in)), filler.hpp
    W = rng.uniform(-scale, scale, size=(K1, K2)).astype(dtype)
    return q, a, W


def make_rerank(Nq, Nc, K, seed=SEED, dtype=np.float32):
    """C4-style inputs: Q^ ~ N(0,1)/sqrt(K) (Nq, K), A^ (Nc, K), W xavier (K, K)."""
    rng = np.random.default_rng(seed)
    Q = (rng.standard_normal((Nq, K)) / np.sqrt(K)).astype(dtype)
    A = (rng.standard_normal((Nc, K)) / np.sqrt(K)).astype(dtype)
    scale = np.sqrt(3.0 / K)
    W = rng.uniform(-scale, scale, size=(K, K)).astype(dtype)
    return Q, A, W
