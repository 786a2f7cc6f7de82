"""Pin the oracle restatement to the REFERENCE'S OWN function bodies, live, on fresh inputs.

Only runs where /root/reference exists (the build container); skipped on the GPU box, where the
committed fixtures of tests/golden/ stand in (tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import reference_shims as RS
from oracle import vtc_oracle as O
from vtc_b200.synthetic import make_batch_pair, make_cam_inputs, make_retrieval_pair

pytestmark = pytest.mark.skipif(not RS.reference_available(), reason="needs /root/reference")


def test_clip_loss_is_the_reference_function():
    loss_mod = RS.ref_loss_module()
    for seed, b, D, s in ((1, 50, 512, 100.0), (2, 128, 512, 14.3), (3, 7, 32, 5.0)):
        vis, txt = make_batch_pair(b, D, seed=seed)
        sim = O.sim_matrix(vis, txt, torch.tensor(s))
        want = loss_mod.clip_loss((vis, txt, sim), {"ignored": True})
        assert torch.equal(O.clip_loss(sim), want)


def test_recall_at_k_is_the_reference_body():
    metric_mod = RS.ref_metric_module()
    for seed, n, D in ((1, 500, 128), (2, 333, 512)):
        T, V = make_retrieval_pair(n, n, D, sigma=3.0, seed=seed)
        for k_vals in ([1, 5, 10], [1, 10], 5):
            m = metric_mod.RecallAtK("a", "b", k_vals)
            want = m.compute(V.numpy(), T.numpy())
            got = O.recall_at_k(V.numpy(), T.numpy(), m.k_vals)
            assert [(int(k), float(r)) for k, r in want] == [(int(k), float(r)) for k, r in got]
            # and the rank definition of this repo gives the same R@k
            r0 = O.rank0_exact(T, V)
            assert [float(r) for _, r in O.recall_from_ranks(r0, m.k_vals)] == [float(r) for _, r in want]


def test_reference_update_result_protocol_keys():
    metric_mod = RS.ref_metric_module()
    T, V = make_retrieval_pair(120, 120, 64, sigma=2.0, seed=5)
    m = metric_mod.RecallAtK("visual", "titles", [1, 10])
    m.writer = None
    for s in range(0, 120, 40):
        m.update(0.0, (V[s:s + 40], T[s:s + 40]), {})
    res = m.result()
    assert set(res) == {"titles_from_visual-recall_at_1", "titles_from_visual-recall_at_10",
                        "visual_from_titles-recall_at_1", "visual_from_titles-recall_at_10"}
    want = dict(O.recall_at_k(V.numpy(), T.numpy(), [1, 10]))
    assert res["titles_from_visual-recall_at_1"] == want[1]


def test_compute_recall_is_the_reference_function():
    reval = RS.ref_retrieval_evaluation_module()
    T, V = make_retrieval_pair(400, 400, 256, sigma=4.0, seed=9)
    want = reval.compute_recall(V, T.unsqueeze(1), split="val", dataset_name="X")
    got = O.compute_recall(V, T.unsqueeze(1), split="val", dataset_name="X")
    assert list(want.index) == list(got.index) and list(want.columns) == list(got.columns)
    np.testing.assert_array_equal(want.values, got.values)


@pytest.mark.parametrize("avg", [True, False])
def test_adapt_feature_is_the_reference_method(avg):
    b, nc, D, layers, heads = 12, 5, 128, 2, 4
    params = O.make_cam_params(D, layers, heads, seed=3, rerandomise=True)
    flw = torch.randn(D, D, generator=torch.Generator().manual_seed(1)) / D ** 0.5
    cam = RS.make_ref_cam(D, layers, heads, params=params, init_from_avg=avg, final_linear_weight=flw)
    main, aux = make_cam_inputs(b, nc, D, seed=4)
    with torch.no_grad():
        want = cam._adapt_feature(main, aux)
    got = O.adapt_feature(main, aux, params, layers, heads, init_from_avg=avg, final_linear_weight=flw)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)


def test_encode_with_comments_routing_matches_reference():
    """_encode_with_comments (model/model.py:216-266) in eval mode with precomputed comment
    features: text / image / skip branches."""
    b, nc, D, layers, heads = 6, 3, 64, 2, 2
    params = O.make_cam_params(D, layers, heads, seed=8, rerandomise=True)
    main, aux = make_cam_inputs(b, nc, D, seed=2)
    vis = torch.randn(b, D, generator=torch.Generator().manual_seed(5))
    for branch in ("text", "image", "skip"):
        cam = RS.make_ref_cam(D, layers, heads, params=params, branch_to_adapt_val=branch)
        cam._load_comment_features = lambda comments: comments  # encoder is out of scope
        with torch.no_grad():
            fv, ft = cam._encode_with_comments(vis, main, aux)
        if branch == "text":
            torch.testing.assert_close(fv, O.normalize(vis))
            torch.testing.assert_close(ft, O.normalize(O.adapt_feature(main, aux, params, layers, heads)),
                                       rtol=1e-5, atol=1e-6)
        elif branch == "image":
            torch.testing.assert_close(ft, O.normalize(main))
            torch.testing.assert_close(fv, O.normalize(O.adapt_feature(vis, aux, params, layers, heads)),
                                       rtol=1e-5, atol=1e-6)
        else:
            torch.testing.assert_close(fv, O.normalize(vis))
            torch.testing.assert_close(ft, O.normalize(main))


def test_constructor_and_call_signatures_match_the_reference():
    """The drop-in boundary is name lookup + keyword construction (train.py:67,85-89): the
    reference's parameter names must be accepted, in the reference's order, by the product."""
    import inspect

    from vtc_b200.evaluation import retrieval_evaluation as our_reval
    from vtc_b200.model import loss as our_loss
    from vtc_b200.model import metric as our_metric
    from vtc_b200.model import model as our_model

    ref_model, ref_metric = RS.ref_model_module(), RS.ref_metric_module()
    ref_loss, ref_reval = RS.ref_loss_module(), RS.ref_retrieval_evaluation_module()

    def leading(fn):
        return [p.name for p in inspect.signature(fn).parameters.values()
                if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]

    def defaults(fn):
        return {p.name: p.default for p in inspect.signature(fn).parameters.values()
                if p.default is not p.empty}

    for cls in ("PretrainedCLIP", "PretrainedCLIP_finaltf", "PretrainedCLIP_TimeSformer",
                "PretrainedCLIP_TimeSformer_finaltf"):
        ref, ours = getattr(ref_model, cls).__init__, getattr(our_model, cls).__init__
        assert leading(ours)[:len(leading(ref))] == leading(ref), cls
        rd, od = defaults(ref), defaults(ours)
        assert all(od[k] == v for k, v in rd.items()), cls
        assert leading(getattr(our_model, cls).forward) == leading(getattr(ref_model, cls).forward), cls
    for name in ("_adapt_feature", "_encode_with_comments"):
        ours_m, ref_m = getattr(our_model.PretrainedCLIPBase, name), getattr(ref_model.PretrainedCLIPBase, name)
        assert leading(ours_m) == leading(ref_m), name
    assert leading(our_metric.RecallAtK.__init__)[:4] == leading(ref_metric.RecallAtK.__init__)
    for meth in ("update", "compute", "reset", "result", "avg", "set_writer"):
        assert leading(getattr(our_metric.RecallAtK, meth)) == leading(getattr(ref_metric.RecallAtK, meth)), meth
    # the trainer imports these two from the same module and drives them through this API
    # (trainer/trainer.py:9,49-54,78-81)
    for cls in ("MetricTracker", "BaseMetric", "ScalarPerBatchMetric", "LossMetric"):
        rc, oc = getattr(ref_metric, cls), getattr(our_metric, cls)
        assert leading(oc.__init__) == leading(rc.__init__), cls
        ref_api = {n for n, v in vars(rc).items() if callable(v) and not n.startswith("_")}
        assert ref_api <= {n for n in dir(oc) if callable(getattr(oc, n))}, (cls, ref_api)
        for meth in ref_api:
            assert leading(getattr(oc, meth)) == leading(getattr(rc, meth)), (cls, meth)
    assert leading(our_loss.clip_loss)[:2] == leading(ref_loss.clip_loss)
    assert leading(our_reval.compute_recall)[:4] == leading(ref_reval.compute_recall)
    assert defaults(ref_reval.compute_recall).items() <= defaults(our_reval.compute_recall).items()


@pytest.mark.parametrize("random_masking", [False, True])
def test_mask_embedding_is_trained_like_in_the_reference(random_masking, monkeypatch):
    """model/model.py:212 and :243-246 route `mask_embedding` into the comment features through
    autograd-visible operations, so the parameter receives gradients.  The glue of
    `_encode_with_comments` is compared with the reference's own method, live, with the CAM itself
    (a CUDA op here) replaced on both sides by the same differentiable stand-in."""
    import vtc_b200.model.model as our_model
    from vtc_b200.model import PretrainedCLIP_finaltf

    # the final normalisation (:263-264) is a CUDA kernel on our side: same formula in torch here
    monkeypatch.setattr(our_model, "normalize", O.normalize)
    b, nc, D = 6, 4, 32
    g = torch.Generator().manual_seed(3)
    vis, title = torch.randn(b, D, generator=g), torch.randn(b, D, generator=g)
    comm = torch.randn(b, nc, D, generator=g)
    empty = torch.rand(b, nc, generator=g) < 0.3
    me = torch.randn(1, D, generator=g)

    def stand_in(main, aux):  # differentiable in both arguments, like the CAM
        aux = torch.stack(list(aux), 0) if not isinstance(aux, torch.Tensor) else aux
        return main + aux.mean(0) * 0.5 + (aux ** 2).sum(0) * 0.1

    ref = RS.make_ref_cam(D, 1, 2, mask_embedding=me)
    ref.random_comment_masking = random_masking
    ref._adapt_feature = stand_in

    def ref_load(comments):  # model/model.py:207-214 without the text encoder (out of scope)
        feats, mask = comments
        feats = feats.clone().float()
        feats[mask] = ref.mask_embedding
        return feats.permute(1, 0, 2)

    ref._load_comment_features = ref_load
    ours = PretrainedCLIP_finaltf(D, n_layers=1, n_heads=2, random_comment_masking=random_masking)
    with torch.no_grad():
        ours.mask_embedding.copy_(me)
    ours._adapt_feature = stand_in
    outs = []
    for m in (ref, ours):
        m.train()
        torch.manual_seed(11)  # the Bernoulli comment masks come from the global generator
        fv, ft = m._encode_with_comments(vis, title, (comm, empty))
        (fv.sum() + (ft * torch.arange(D)).sum()).backward()
        outs.append((fv.detach(), ft.detach(), m.mask_embedding.grad.clone()))
    torch.testing.assert_close(outs[1][0], outs[0][0])
    torch.testing.assert_close(outs[1][1], outs[0][1])
    assert outs[0][2].abs().max() > 0
    torch.testing.assert_close(outs[1][2], outs[0][2])


def test_averaging_fusion_forward_glue_matches_reference(monkeypatch):
    """PretrainedCLIP.forward with comment_fusion="averaging" (model/model.py:326-371) through a
    stand-in backbone and TOKEN-ID inputs: the comments are averaged in as the encoder sees them
    (no mask_embedding in this class), both outputs are normalised, sim = exp(logit_scale) * v t^T.
    The two CUDA ops on our side (normalize, the fused mean + normalise read-out) are replaced by
    their torch formulas; everything else is the code that runs on the box."""
    import torch.nn as nn

    import vtc_b200.model.model as our_model

    D, b, nc, ntok = 16, 5, 3, 7

    class FakeClip(nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(2)
            self.emb = nn.Parameter(torch.randn(50000, D, generator=g))
            self.proj = nn.Parameter(torch.randn(3 * 4 * 4, D, generator=g))
            self.logit_scale = nn.Parameter(torch.tensor(2.5))
            self.ln_final = nn.LayerNorm(D)

        def encode_text(self, tok):
            return self.emb[tok].mean(1)

        def encode_image(self, x):
            return x.flatten(1) @ self.proj

    backbone = FakeClip()
    mm = RS.ref_model_module()

    class RefClip(mm.PretrainedCLIP):
        def __init__(self):
            nn.Module.__init__(self)
            self.model = backbone
            self.feature_dim = D
            self.residual_activation = None
            self.comment_fusion = "averaging"

    monkeypatch.setattr(our_model, "normalize", O.normalize)
    monkeypatch.setattr(our_model._UniformReadout, "apply",
                        staticmethod(lambda stacked: O.normalize(stacked.mean(0))))
    ours = our_model.PretrainedCLIP("ViT-B/32", comment_fusion="averaging", backbone=backbone,
                                    lazy_sim=False)
    monkeypatch.setattr(our_model.LazySim, "materialize",
                        lambda self: self.scale * self.feats_a @ self.feats_b.t())
    g = torch.Generator().manual_seed(9)
    vis = torch.randn(b, 2, 3, 4, 4, generator=g)              # [b, t, c, h, w]: mean over time
    title = torch.randint(0, 49000, (b, ntok), generator=g)
    comments = torch.randint(0, 49000, (b, nc, ntok), generator=g)
    comments[1, 2, 1] = 49407                                   # an empty comment
    with torch.no_grad():
        want = RefClip().eval()(vis, title, comments)
        got = ours.eval()(vis, title, comments)
    for g_, w_ in zip(got, want):
        torch.testing.assert_close(g_, w_, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("branch", ["text", "image", "skip"])
def test_finaltf_forward_glue_with_token_ids_matches_reference(monkeypatch, branch):
    """PretrainedCLIP_finaltf.forward (model/model.py:456-480) with TOKEN-ID titles / comments and
    raw frames through a stand-in backbone: the reference's own _load_comment_features (:207-214,
    empty comment = end-of-text token 49407 in position 1 -> mask_embedding) and
    _encode_with_comments run unmodified; the CAM is the same differentiable stand-in on both sides."""
    import torch.nn as nn

    import vtc_b200.model.model as our_model

    D, b, nc, ntok = 16, 5, 3, 7

    class FakeClip(nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(4)
            self.emb = nn.Parameter(torch.randn(50000, D, generator=g))
            self.proj = nn.Parameter(torch.randn(3 * 4 * 4, D, generator=g))
            self.logit_scale = nn.Parameter(torch.tensor(2.0))
            self.ln_final = nn.LayerNorm(D)

        def encode_text(self, tok):
            return self.emb[tok].mean(1)

        def encode_image(self, x):
            return x.flatten(1) @ self.proj

    backbone = FakeClip()
    me = torch.randn(1, D, generator=torch.Generator().manual_seed(6))

    def stand_in(main, aux):
        aux = torch.stack(list(aux), 0) if not isinstance(aux, torch.Tensor) else aux
        return main * 0.7 + aux.mean(0)

    mm = RS.ref_model_module()

    class RefFinal(mm.PretrainedCLIP_finaltf):
        def __init__(self):
            nn.Module.__init__(self)
            self.model = backbone
            self.feature_dim = D
            self.mask_embedding = nn.Parameter(me.clone())
            self.random_comment_masking = False
            self.branch_to_adapt = self.branch_to_adapt_val = branch
            self.init_audio_model = False

    ref = RefFinal().eval()
    ref._adapt_feature = stand_in
    monkeypatch.setattr(our_model, "normalize", O.normalize)
    monkeypatch.setattr(our_model.LazySim, "materialize",
                        lambda self: self.scale * self.feats_a @ self.feats_b.t())
    ours = our_model.PretrainedCLIP_finaltf("ViT-B/32", branch_to_adapt=branch,
                                            branch_to_adapt_val=branch, n_layers=1, n_heads=2,
                                            backbone=backbone, lazy_sim=False).eval()
    with torch.no_grad():
        ours.mask_embedding.copy_(me)
    ours._adapt_feature = stand_in
    g = torch.Generator().manual_seed(9)
    vis = torch.randn(b, 3, 4, 4, generator=g)                  # [b, c, h, w]
    title = torch.randint(0, 49000, (b, ntok), generator=g)
    comments = torch.randint(0, 49000, (b, nc, ntok), generator=g)
    comments[1, 2, 1] = 49407
    comments[3, 0, 1] = 49407
    with torch.no_grad():
        want = ref(vis, title, comments)
        got = ours(vis, title, comments)
    for g_, w_ in zip(got, want):
        torch.testing.assert_close(g_, w_, rtol=1e-5, atol=1e-6)


def test_committed_goldens_are_reproduced_by_the_generator(tmp_path, monkeypatch):
    """tests/golden/*.npz are what the GPU box checks against (the reference does not exist there),
    so they must be exactly what tests/golden/generate_golden.py -- i.e. the reference's own code --
    produces today: regenerate into a scratch directory and compare every array."""
    import importlib.util
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location(
        "generate_golden", os.path.join(here, "golden", "generate_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    monkeypatch.setattr(gen, "OUT", str(tmp_path))
    gen.main()
    made = sorted(f for f in os.listdir(tmp_path) if f.endswith(".npz"))
    committed = sorted(f for f in os.listdir(os.path.join(here, "golden")) if f.endswith(".npz"))
    assert made == committed
    for name in made:
        new = np.load(os.path.join(tmp_path, name), allow_pickle=True)
        old = np.load(os.path.join(here, "golden", name), allow_pickle=True)
        assert sorted(new.files) == sorted(old.files), name
        for key in new.files:
            a, b = new[key], old[key]
            if a.dtype.kind in "fc":
                np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7, err_msg=f"{name}:{key}")
            else:
                np.testing.assert_array_equal(a, b, err_msg=f"{name}:{key}")


def test_reference_metric_module_runs_unmodified_on_the_faiss_compat_binding(monkeypatch):
    """The other way to switch (vtc_b200/faiss_compat.py): the reference's model/metric.py, byte for
    byte, with its `import faiss` bound to the B200 library.  It imports, RecallAtK constructs and
    accumulates exactly as before (:103-135), and compute() reaches GpuIndexFlatL2 -- which, in this
    container without a GPU, must fail loudly instead of falling back to a CPU search."""
    import collections
    import importlib.util
    import os
    import sys

    from vtc_b200 import faiss_compat
    from vtc_b200._ffi import VtcError

    RS.install()   # (collections.Iterable for model/metric.py:106)
    assert hasattr(collections, "Iterable")
    monkeypatch.setitem(sys.modules, "faiss", faiss_compat)
    spec = importlib.util.spec_from_file_location(
        "reference_metric_on_vtc_b200", os.path.join(RS.REFERENCE_ROOT, "model", "metric.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.faiss is faiss_compat
    m = mod.RecallAtK("videos", "titles", [1, 5, 10])
    assert isinstance(m.knn_config, faiss_compat.GpuIndexFlatConfig) and m.knn_config.useFloat16 is False
    T, V = make_retrieval_pair(64, 64, 32, sigma=2.0, seed=4)
    m.update(None, (V[:40], T[:40]), {})
    m.update(None, (V[40:], T[40:]), {})
    assert m.insert_index == 64 and m.knn_config.device is None   # CPU tensors: fa.device.index
    if not torch.cuda.is_available():
        with pytest.raises(VtcError):
            m.compute(V.numpy(), T.numpy())
    else:
        got = m.compute(V.numpy(), T.numpy())
        want = O.recall_at_k(V.numpy(), T.numpy(), [1, 5, 10])
        assert [(int(k), float(r)) for k, r in got] == [(int(k), float(r)) for k, r in want]
