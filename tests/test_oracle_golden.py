"""The CPU oracle (oracle/vtc_oracle.py + .c) against the committed golden fixtures, which were
produced by executing the REFERENCE'S OWN function bodies (tests/golden/generate_golden.py)."""
import hashlib

import numpy as np
import torch

from oracle import vtc_oracle as O
from vtc_b200.synthetic import make_batch_pair, make_cam_inputs, make_retrieval_pair


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_c1_inputs_are_the_ones_the_golden_was_made_from(golden):
    g = golden("retrieval_c1.npz")
    T, V = make_retrieval_pair(1000, 1000, 512, sigma=6.0, seed=1023)
    assert sha(T.numpy(), V.numpy()) == str(g["input_sha"])
    Tm, Vm = make_retrieval_pair(1000, 1000, 512, seed=1023, mixed=True)
    assert sha(Tm.numpy(), Vm.numpy()) == str(g["input_sha_mixed"])


def test_c1_compute_recall_matches_reference_dataframe(golden):
    g = golden("retrieval_c1.npz")
    for mixed, key in ((False, "df_values"), (True, "df_values_mixed")):
        T, V = make_retrieval_pair(1000, 1000, 512, sigma=None if mixed else 6.0, seed=1023, mixed=mixed)
        df = O.compute_recall(V, T.unsqueeze(1))
        assert list(df.index) == list(g["df_index"])
        assert list(df.columns) == list(g["df_columns"])
        np.testing.assert_array_equal(df.values, g[key])


def test_c1_rank_definition_agrees_with_reference_recall(golden):
    """R@k from rank0 (this repo's definition) == the reference's 'gt in top-k list'."""
    g = golden("retrieval_c1.npz")
    for sfx, col in (("", "df_values"), ("_mixed", "df_values_mixed")):
        for direction, c in (("rank_v2t", 0), ("rank_t2v", 1)):
            ranks = g[direction + sfx]
            r = np.array([x for _, x in O.recall_from_ranks(ranks, [1, 5, 10])]) * 100.0
            np.testing.assert_allclose(r, g[col][:, c], rtol=0, atol=1e-12)


def test_c1_rank_oracle_reproduces_golden_ranks(golden):
    g = golden("retrieval_c1.npz")
    T, V = make_retrieval_pair(1000, 1000, 512, sigma=6.0, seed=1023)
    np.testing.assert_array_equal(O.rank0_exact(T, V), g["rank_t2v"])
    np.testing.assert_array_equal(O.rank0_exact(V, T), g["rank_v2t"])


def test_small_adversarial_case(golden):
    g = golden("retrieval_small.npz")
    Q, G = g["queries"], g["gallery"]
    k_vals = list(g["k_vals"])
    got = np.array([r for _, r in O.recall_at_k(G, Q, k_vals)])
    np.testing.assert_array_equal(got, g["recall_q_from_g"])
    got = np.array([r for _, r in O.recall_at_k(Q, G, k_vals)])
    np.testing.assert_array_equal(got, g["recall_g_from_q"])
    # fp64 ranks vs the reference's top-11 lists: rank < 11  <=>  t in list, position == rank
    rank0 = O.rank0_exact(Q, G)
    np.testing.assert_array_equal(rank0, g["rank0"])
    I = g["top11_idx"]
    for t in range(Q.shape[0]):
        if np.isnan(Q[t]).any():
            assert rank0[t] == G.shape[0]  # NaN query: never retrieved
            continue
        pos = np.where(I[t] == t)[0]
        if rank0[t] < 11:
            assert len(pos) == 1 and pos[0] == rank0[t], (t, rank0[t], I[t])
        else:
            assert len(pos) == 0
    # exact top-k agrees with the fp32 stand-in wherever scores are not within fp32 rounding
    vals, idx = O.topk_exact(Q, G, 11)
    ok = ~np.isnan(Q).any(1)
    agree = (idx[ok] == I[ok]).mean()
    assert agree > 0.99


def test_clip_loss_golden(golden):
    g = golden("clip_loss.npz")
    for name in ("c2_s100", "c2_s14", "small", "ragged"):
        b, D, s = g[name + "_cfg"]
        vis, txt = make_batch_pair(int(b), int(D), seed=1023)
        assert sha(vis.numpy(), txt.numpy()) == str(g[name + "_sha"])
        sim = O.sim_matrix(vis, txt, torch.tensor(float(s)))
        loss = O.clip_loss(sim)
        np.testing.assert_allclose(loss.item(), g[name + "_loss"], rtol=1e-6)
        p64 = O.clip_loss_parts64(vis, txt, float(s))
        np.testing.assert_allclose(p64["loss"], g[name + "_loss"], rtol=2e-5)


def test_cam_golden(golden):
    g = golden("cam.npz")
    for name in ("c2_init", "c2_rand", "c2_linear", "small"):
        b, nc, D, layers, heads, rerand, avg = [int(x) for x in g[name + "_cfg"]]
        params = O.make_cam_params(D, layers, heads, seed=1023, rerandomise=bool(rerand))
        main, aux = make_cam_inputs(b, nc, D, seed=1023)
        assert sha(main.numpy(), aux.numpy(), *[params[k].numpy() for k in sorted(params)]) == \
            str(g[name + "_sha"])
        gen = torch.Generator().manual_seed(99)
        flw = torch.randn(D, D, generator=gen) / D ** 0.5
        out = O.adapt_feature(main, aux, params, layers, heads, init_from_avg=bool(avg),
                              final_linear_weight=flw)
        want = g[name + "_adapted"]
        np.testing.assert_allclose(out[:want.shape[0]].numpy(), want, rtol=1e-4, atol=2e-6)
        x = O.normalize(torch.stack([main, *aux], 0))
        tfm = O.transformer(x, params, layers, heads)
        np.testing.assert_allclose(tfm[0, :8].numpy(), g[name + "_tfm_tok0"], rtol=1e-4, atol=2e-6)


def test_cam_residual_activations_golden(golden):
    """model/model.py:30-77 variants against what the reference's own table produced."""
    g = golden("cam.npz")
    b, nc, D, layers, heads = 8, 3, 64, 2, 2
    params = O.make_cam_params(D, layers, heads, seed=1023, rerandomise=True)
    main, aux = make_cam_inputs(b, nc, D, seed=1023)
    bn = (torch.from_numpy(g["act_running_mean"]), torch.from_numpy(g["act_running_var"]), 1e-5)
    for act in ("normalize", "squash", "squash10", "squash1p5", "tanh", "sub_mean", "bn"):
        out = O.adapt_feature(main, aux, params, layers, heads, residual_activation=act, bn_state=bn)
        np.testing.assert_allclose(out.numpy(), g["act_" + act], rtol=1e-4, atol=2e-6)


def test_cam_closed_form_at_init():
    """reference zero-inits make the transformer an identity."""
    params = O.make_cam_params(512, 2, 8, seed=1023)
    main, aux = make_cam_inputs(32, 5, 512)
    out = O.adapt_feature(main, aux, params, 2, 8)
    np.testing.assert_allclose(out.numpy(), O.cam_closed_form_at_init(main, aux).numpy(),
                               rtol=1e-5, atol=1e-6)


def test_rank_oracle_properties():
    T, V = make_retrieval_pair(300, 500, 64, sigma=2.0, seed=5)
    r = O.rank0_exact(T, V)
    # sharding the gallery: ranks are additive over chunks when gt scores are global
    full = O.scores64(T, V)
    d0 = full[np.arange(300), np.arange(300)]
    brute = np.array([(full[t] < d0[t]).sum() + ((full[t] == d0[t]) & (np.arange(500) < t)).sum()
                      for t in range(300)])
    np.testing.assert_array_equal(r, brute)
    # top-k consistency
    vals, idx = O.topk_exact(T, V, 7)
    for t in range(300):
        order = np.lexsort((np.arange(500), full[t]))[:7]
        np.testing.assert_array_equal(idx[t], order)
    assert O.medr(np.array([0, 0, 3, 9])) == 2.5


def test_eval_tail_matches_reference_semantics():
    vids = [torch.randn(n, 8) for n in (1, 3, 2)]
    caps = [torch.randn(n, 8) for n in (1, 2, 1)]
    v, c = O.eval_tail(vids, caps)
    assert v.shape == (3, 8) and c.shape == (3, 2, 8)
    assert torch.isinf(c[0, 1]).all() and (c[0, 1] < 0).all()
    torch.testing.assert_close(v[1], vids[1].mean(0))


def test_c_oracle_against_a_pure_python_restatement_of_the_canonical_arithmetic():
    """The C oracle is the anchor every CUDA rank / top-k is compared with, so it is itself pinned to
    the definition written out in plain Python: dot = fold_k acc + q_k * x_k in k
    order (the product of two floats is exact in a double, so `acc + a * b` IS the fused
    multiply-add), d = sq(x) - 2 dot (L2) or -dot (DOT), rank0 = #{j != gt: d_j < d_gt} +
    #{j < gt: d_j == d_gt}, NaN comparisons false, top-k by (score, index).  Values come from a
    small grid so that exact ties, duplicates, zero rows and a NaN are all present."""
    rng = np.random.default_rng(7)
    grid = np.array([-1.5, -0.5, 0.0, 0.25, 0.5, 1.0, 3.0], dtype=np.float32)
    N, M, D = 23, 41, 9
    Q = rng.choice(grid, size=(N, D)).astype(np.float32)
    G = rng.choice(grid, size=(M, D)).astype(np.float32)
    G[5] = G[30]                      # duplicate rows -> exact ties
    G[7] = 0.0                        # zero row
    Q[:10] = G[:10] + rng.choice(np.array([0.0, 0.25], dtype=np.float32), size=(10, D))
    Q[3, 2] = np.nan
    gt = rng.integers(0, M, size=N)

    def dot(a, b):
        acc = 0.0
        for x, y in zip(a.tolist(), b.tolist()):
            acc = acc + x * y
        return acc

    for metric in (O.METRIC_L2, O.METRIC_DOT):
        d = np.empty((N, M))
        for t in range(N):
            for j in range(M):
                d[t, j] = (dot(G[j], G[j]) - 2.0 * dot(Q[t], G[j])) if metric == O.METRIC_L2 \
                    else -dot(Q[t], G[j])
        np.testing.assert_array_equal(O.scores64(Q, G, metric), d)
        want = np.zeros(N, dtype=np.int64)
        for t in range(N):
            d0 = d[t, gt[t]]
            for j in range(M):
                if j != gt[t] and (d[t, j] < d0 or (d[t, j] == d0 and j < gt[t])):
                    want[t] += 1
        got = O.rank0_exact(Q, G, gt=gt, metric=metric)
        ok = ~np.isnan(d[np.arange(N), gt])   # a NaN score: vtc_rank_finalize's business (rank = M)
        np.testing.assert_array_equal(got[ok], want[ok])
        vals, idx = O.topk_exact(Q, G, 6, metric=metric)
        for t in range(N):
            if np.isnan(d[t]).any():
                continue
            order = sorted(range(M), key=lambda j: (d[t, j], j))[:6]
            np.testing.assert_array_equal(idx[t], order)
            np.testing.assert_array_equal(vals[t], d[t, order])
