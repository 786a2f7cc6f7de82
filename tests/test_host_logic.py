"""CPU tests of the host-side mirror of the reference interface (no GPU work)."""
import numpy as np
import pytest
import torch


def test_reference_symbols_and_signatures_exist():
    import inspect

    from vtc_b200.evaluation.retrieval_evaluation import compute_recall
    from vtc_b200.model.loss import clip_loss
    from vtc_b200.model.metric import RecallAtK

    sig = inspect.signature(compute_recall)
    assert list(sig.parameters)[:4] == ["tensor_v", "tensor_t", "split", "dataset_name"]
    assert sig.parameters["split"].default == "full-test"
    assert sig.parameters["dataset_name"].default == "MSRVTT"
    assert list(inspect.signature(clip_loss).parameters)[:2] == ["input", "meta"]
    m = RecallAtK("visual", "titles", k_vals=5)  # scalar k like the reference default
    assert m.k_vals == [5] and m.name == "recall@k" and m.is_train is False and m.is_val is True
    for meth in ("set_writer", "reset", "update", "avg", "result", "compute"):
        assert callable(getattr(m, meth))
    assert m.avg() is None


def test_checkpoint_parameter_names_match_the_reference():
    """final_transformer.resblocks.{i}.* / final_linear.weight / mask_embedding must survive
    (trainer/base_trainer.py:175-176, train.py:105)."""
    from oracle.vtc_oracle import cam_param_names
    from vtc_b200.model import PretrainedCLIP_finaltf

    m = PretrainedCLIP_finaltf(64, n_layers=2, n_heads=2)
    names = {n for n, _ in m.named_parameters()}
    for n in cam_param_names(2):
        assert "final_transformer." + n in names, n
    assert "final_linear.weight" in names and "mask_embedding" in names
    # the reference's zero-inits (model/model.py:440-452)
    for blk in m.final_transformer.resblocks:
        assert blk.mlp.c_proj.weight.abs().sum() == 0 and blk.mlp.c_proj.bias.abs().sum() == 0
        assert blk.attn.out_proj.weight.abs().sum() == 0
    assert m.final_linear.weight.abs().sum() == 0
    # shapes follow nn.MultiheadAttention / clip.model.Transformer
    sd = m.state_dict()
    assert sd["final_transformer.resblocks.0.attn.in_proj_weight"].shape == (192, 64)
    assert sd["final_transformer.resblocks.1.mlp.c_fc.weight"].shape == (256, 64)


def test_lazy_sim_metadata_needs_no_gpu():
    from vtc_b200.model import LazySim

    a, b = torch.randn(5, 8), torch.randn(7, 8)
    s = LazySim(a, b, 100.0)
    assert tuple(s.shape) == (5, 7) and s.size(0) == 5 and s.dim() == 2 and len(s) == 5
    assert s.dtype == torch.float32 and s.device == a.device


def test_multi_caption_input_is_rejected_clearly():
    from vtc_b200.evaluation.retrieval_evaluation import _squeeze_text

    with pytest.raises(ValueError, match="one text per video"):
        _squeeze_text(torch.zeros(10, 20, 8))
    assert _squeeze_text(torch.zeros(10, 1, 8)).shape == (10, 8)
    assert _squeeze_text(torch.zeros(1, 1, 8)).shape == (1, 8)


def test_shard_bounds_partition():
    from vtc_b200.parallel import shard_bounds

    for n in (0, 1, 7, 100, 100003):
        for w in (1, 2, 3, 8):
            cuts = [shard_bounds(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_recipe_is_calibrated():
    from vtc_b200.synthetic import make_retrieval_pair

    T, V = make_retrieval_pair(1000, 1000, 512)
    np.testing.assert_allclose(T.norm(dim=-1).numpy(), 1.0, rtol=1e-5)
    cos = (T * V).sum(1)
    assert abs(cos.mean().item() - 1 / np.sqrt(1 + 36)) < 0.01


def test_cached_embedding_reader_and_alignment(tmp_path):
    """The reference's cached-feature format (scripts/get_clip_vit_embeddings.py:72-78)."""
    from vtc_b200.data import align_by_id, load_cached_embeddings

    g = torch.Generator().manual_seed(0)
    ids_a = torch.tensor([7, 3, 11, 5, 2], dtype=torch.int64)
    ids_b = torch.tensor([5, 7, 2, 99, 3], dtype=torch.int64)
    ea, eb = torch.randn(5, 8, generator=g), torch.randn(5, 8, generator=g)
    pa, pb = tmp_path / "a.pth", tmp_path / "b.pth"
    torch.save({"reddit_ids": ids_a, "embeddings": ea}, pa)
    torch.save({"reddit_ids": ids_b, "embeddings": eb}, pb)
    ia, xa = load_cached_embeddings(str(pa), pin=False)
    ib, xb = load_cached_embeddings(str(pb), pin=False)
    assert torch.equal(ia, ids_a) and torch.equal(xa, ea)
    ids, va, vb = align_by_id(ia, xa, ib, xb)
    assert ids.tolist() == [7, 3, 5, 2]  # file-a order, shared ids only
    assert torch.equal(va[0], ea[0]) and torch.equal(vb[0], eb[1])
    assert torch.equal(va[3], ea[4]) and torch.equal(vb[3], eb[2])
    ids2, _, _ = align_by_id(ia, xa, ib, xb, order=[2, 7])
    assert ids2.tolist() == [2, 7]
    with pytest.raises(KeyError):
        align_by_id(ia, xa, ib, xb, order=[11])
    torch.save({"reddit_ids": ids_a.int(), "embeddings": ea}, pa)
    with pytest.raises(AssertionError):
        load_cached_embeddings(str(pa), pin=False)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's CPU arm): one JSON line with the contract's keys,
    no GPU needed, no work on ranks other than 0."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--n", "1500",
           "--m", "1500", "--steps", "1", "--warmup", "0", "--cpu-seconds", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]
    # under torchrun only rank 0 works and prints
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "1"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_eval_tail_vectorised_matches_the_oracle_on_cpu():
    """evaluation/retrieval_evaluation.py:238-260: -inf caption padding (exact) and per-video frame
    mean (fp32 reduction order may differ: allclose), via the device-agnostic core of eval_tail."""
    import torch

    from oracle import vtc_oracle as O
    from vtc_b200.evaluation.retrieval_evaluation import _eval_tail_on

    g = torch.Generator().manual_seed(0)
    vids = [torch.randn(int(n), 32, generator=g) for n in torch.randint(1, 40, (150,), generator=g)]
    caps = [torch.randn(int(n), 32, generator=g) for n in torch.randint(1, 6, (150,), generator=g)]
    v, c = _eval_tail_on(vids, caps, torch.device("cpu"))
    vo, co = O.eval_tail(vids, caps)
    assert torch.equal(c, co)
    assert v.shape == vo.shape and torch.allclose(v, vo, rtol=1e-6, atol=1e-7)
    # the single-frame / single-caption case degenerates to plain stacking
    v1, c1 = _eval_tail_on([x[:1] for x in vids], [x[:1] for x in caps], torch.device("cpu"))
    assert torch.equal(v1, torch.cat([x[:1] for x in vids]))
    assert torch.equal(c1[:, 0], torch.cat([x[:1] for x in caps]))


def test_models_construct_from_the_reference_config_args():
    """`config.init_obj("arch", module_arch)` (train.py:67, utils/parse_config.py:97-112) calls
    the class with the `args` dict of the reference's configs: same keyword names must work."""
    import pytest

    from vtc_b200.model.model import PretrainedCLIP, PretrainedCLIP_finaltf

    args = {"model_type": "ViT-B/32", "branch_to_adapt": "text", "branch_to_adapt_val": "text",
            "n_layers": 2, "n_heads": 8, "init_from_avg": True, "random_comment_masking": False,
            "random_skip_adapter": True}          # configs/pretrained_clip_comments_attention.jsonc:7-17
    m = PretrainedCLIP_finaltf(**args)
    assert m.feature_dim == 512 and len(m.final_transformer.resblocks) == 2
    assert m.branch_to_freeze is False
    assert PretrainedCLIP_finaltf("ViT-L/14").feature_dim == 768
    assert PretrainedCLIP_finaltf(64, n_layers=1, n_heads=2).feature_dim == 64   # int shorthand
    # freeze="finaltf" (experiments freeze the adapter, model/model.py:290-299)
    f = PretrainedCLIP_finaltf(model_type="ViT-B/32", freeze="finaltf")
    assert f.branch_to_freeze == "finaltf"
    assert not any(p.requires_grad for p in f.final_transformer.parameters())
    assert not f.final_linear.weight.requires_grad and not f.mask_embedding.requires_grad
    with pytest.raises(Exception, match="Unknown branch_to_freeze"):
        PretrainedCLIP_finaltf(freeze="bogus")
    with pytest.raises(NotImplementedError):
        PretrainedCLIP_finaltf(init_audio_model=True)
    p = PretrainedCLIP(model_type="ViT-B/32", freeze=False, residual_activation=None,
                       comment_fusion="averaging")                  # model/model.py:309-315
    assert p.feature_dim == 512 and p.comment_fusion == "averaging"
    with pytest.raises(ValueError):
        PretrainedCLIP("no-such-clip")
    # the TimeSformer variants share the hot path; their towers come in as `backbone`
    from vtc_b200.model.model import PretrainedCLIP_TimeSformer, PretrainedCLIP_TimeSformer_finaltf

    t = PretrainedCLIP_TimeSformer_finaltf(model_type="ViT-B/32", freeze=False, visual_device=None,
                                           n_layers=2, n_heads=8)     # configs/..timesformer..jsonc
    assert t.feature_dim == 512 and t.multigpu is False
    assert {n for n, _ in t.named_parameters()} == {n for n, _ in m.named_parameters()}
    assert PretrainedCLIP_TimeSformer("ViT-B/32").comment_fusion is None


def test_clock_sampler_prefers_samples_inside_the_timed_region():
    """bench.py's `clocks` object: nvidia-smi lines are time-stamped on arrival, only those inside
    the timed region count, with a labelled fall-back when the region was too short to catch one."""
    import os
    import sys
    import time

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    capped = "0, 1700, 1965, 950.5, 0x4, Not Active, Not Active, Not Active, Active"
    idle = "0, 1965, 1965, 300.0, 0x0, Not Active, Not Active, Not Active, Not Active"
    now = time.perf_counter()
    cs = bench.ClockSampler(0)
    cs.lines = [(now - 1.0, idle), (now - 0.5, idle)]
    cs.window(now - 0.2, now)
    out = cs.summarise()
    assert out["sm_mhz"] == 1965.0 and out["reasons"] == [] and out["window"].startswith("warm-up")
    cs.lines += [(now - 0.1, capped), (now - 0.05, capped), (now + 5.0, idle)]
    out = cs.summarise()
    assert out["window"] == "timed blocks" and out["samples"] == 2
    assert out["sm_mhz"] == 1700.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    assert bench.ClockSampler(0).stop()["reasons"] == ["nvidia-smi unavailable"]


def test_pipelined_staging_chunk_bounds():
    """RecallAtK's host-staging chunk bounds (the H2D pipeline behind RecallAtK.compute, which
    replaces the reference's faiss add/search on host arrays, model/metric.py:140-146): equal
    chunks, monotone, covering [0, n]."""
    from vtc_b200.model.metric import RecallAtK

    assert RecallAtK._pipeline_bounds_2d(100, 4) == [0, 25, 50, 75, 100]
    for n in (0, 1, 7, 1000, 100_000):
        for c in (3, 6, 10):
            b = RecallAtK._pipeline_bounds_2d(n, c)
            assert len(b) == c + 1 and b[0] == 0 and b[-1] == n
            assert all(x <= y for x, y in zip(b[:-1], b[1:]))


def test_metric_tracker_and_loss_metric_follow_the_trainer_protocol():
    """trainer/trainer.py:49-54,78-81,118-121: MetricTracker(*metrics) + add_metric(LossMetric()) +
    set_writer(writer); update(loss.item(), output, meta); result() merges the metrics' dicts;
    avg() maps every metric name to its running average (None for RecallAtK)."""
    from vtc_b200.model.metric import LossMetric, MetricTracker, RecallAtK

    class Writer:
        def __init__(self):
            self.scalars = []

        def add_scalar(self, name, value):
            self.scalars.append((name, value))

    rec = RecallAtK("visual", "titles", k_vals=[1])
    tr = MetricTracker(*[m for m in [rec] if m.is_val])
    tr.add_metric(LossMetric())
    w = Writer()
    tr.set_writer(w)
    assert list(tr.metrics) == ["recall@k", "loss"] and rec.writer is w
    loss_metric = tr.metrics["loss"]
    for v in (2.0, 4.0, 6.0):
        loss_metric.update(v, None, None)
    assert tr.avg() == {"recall@k": None, "loss": 4.0}
    assert loss_metric.result() == {"loss": 4.0}
    assert w.scalars == [("loss", 2.0), ("loss", 4.0), ("loss", 6.0)]
    tr.reset()
    assert loss_metric.avg() == 0 and rec.insert_index == 0


def test_faiss_compat_surface_and_no_cpu_fallback():
    """vtc_b200.faiss_compat exposes the three names model/metric.py:112-113,139-146 uses; without a
    CUDA device the index refuses to exist (no CPU search behind the reference's back)."""
    from vtc_b200 import faiss_compat as faiss
    from vtc_b200._ffi import VtcError

    cfg = faiss.GpuIndexFlatConfig()
    assert cfg.useFloat16 is False and cfg.device == 0
    cfg.useFloat16, cfg.device = False, None          # what RecallAtK.update leaves for CPU tensors
    faiss.StandardGpuResources()
    for name in ("add", "search", "reset", "ntotal"):
        assert hasattr(faiss.GpuIndexFlatL2, name)
    if not torch.cuda.is_available():
        with pytest.raises(VtcError):
            faiss.GpuIndexFlatL2(faiss.StandardGpuResources(), 512, cfg)
    with pytest.raises(ValueError):
        faiss.GpuIndexFlatL2(None, 0, cfg)
