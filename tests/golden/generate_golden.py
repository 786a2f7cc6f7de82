"""Generate the committed golden fixtures by executing the REFERENCE'S OWN function bodies.

Run in the build container only (needs /root/reference):

    python tests/golden/generate_golden.py

The reference is imported through the in-memory stand-ins of oracle/reference_shims.py (fake
`faiss` = exact fp32 L2 on CPU, stand-in `clip.model.Transformer`); see SURVEY.md Appendix A.
Outputs (small .npz files next to this script) are what tests/test_oracle_golden.py checks the
CPU restatement (oracle/vtc_oracle.py) against, and what the `-m gpu` parity tests check the
CUDA path against on the GPU box, where /root/reference does not exist.

Inputs come from vtc_b200.synthetic (seeded torch CPU RNG); each fixture stores a sha256 of the
exact input bytes so that a silent RNG change is caught instead of producing a vacuous pass.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_shims as RS  # noqa: E402
from oracle import vtc_oracle as O  # noqa: E402
from vtc_b200.synthetic import make_batch_pair, make_cam_inputs, make_retrieval_pair  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def ref_recall(metric_mod, gallery: np.ndarray, queries: np.ndarray, k_vals):
    m = metric_mod.RecallAtK("a", "b", k_vals)
    return np.array([r for _, r in m.compute(gallery, queries)], dtype=np.float64)


def gen_retrieval_c1(metric_mod, reval_mod):
    """BASELINE config 1: 1k text x 1k video, D=512, via the reference's compute_recall."""
    T, V = make_retrieval_pair(1000, 1000, 512, sigma=6.0, seed=1023)
    df = reval_mod.compute_recall(V, T.unsqueeze(1), split="full-test", dataset_name="MSRVTT")
    Tm, Vm = make_retrieval_pair(1000, 1000, 512, seed=1023, mixed=True)
    dfm = reval_mod.compute_recall(Vm, Tm.unsqueeze(1), split="full-test", dataset_name="MSRVTT")
    # definitions this repo adds (not in the reference): rank0 / MedR from the fp64 oracle
    np.savez_compressed(
        os.path.join(OUT, "retrieval_c1.npz"),
        input_sha=sha(T.numpy(), V.numpy()), input_sha_mixed=sha(Tm.numpy(), Vm.numpy()),
        df_values=df.values, df_index=np.array(df.index), df_columns=np.array(df.columns),
        df_values_mixed=dfm.values,
        rank_t2v=O.rank0_exact(T, V), rank_v2t=O.rank0_exact(V, T),
        rank_t2v_mixed=O.rank0_exact(Tm, Vm), rank_v2t_mixed=O.rank0_exact(Vm, Tm),
    )
    print("retrieval_c1:\n", df, "\nmixed:\n", dfm)


def gen_retrieval_small(metric_mod):
    """256 x 256 x 64 with stored inputs, including the adversarial rows of SURVEY.md §8d."""
    T, V = make_retrieval_pair(256, 256, 64, sigma=2.5, seed=7)
    T = T.clone()
    V = V.clone()
    V[10] = V[3]            # exact duplicate gallery rows (ties)
    V[11] = V[3]
    T[3] = V[3]             # query identical to three gallery rows
    V[20] = 0.5 * (V[20] + V[21])   # non-unit gallery row (mean of unit vectors, reval.py:254-259)
    V[30] = 0.0             # zero row
    T[40] = float("nan")    # a NaN query (normalising a zero row, model.py:26-27)
    V[50] = float("-inf")   # a -inf padding row (retrieval_evaluation.py:238-252), also a ground truth
    V[51, :3] = float("inf")   # a row with +inf elements: +-inf / NaN scores depending on the query
    k_vals = [1, 5, 10]
    m = metric_mod.RecallAtK("a", "b", k_vals)
    r_ab = np.array([r for _, r in m.compute(V.numpy(), T.numpy())])
    r_ba = np.array([r for _, r in m.compute(T.numpy(), V.numpy())])
    # the top-11 index lists the reference inspects (model/metric.py:144-146)
    import faiss  # the shim

    idx = faiss.GpuIndexFlatL2(None, 64, None)
    idx.add(V.numpy())
    D11, I11 = idx.search(T.numpy(), 11)
    np.savez_compressed(
        os.path.join(OUT, "retrieval_small.npz"),
        queries=T.numpy(), gallery=V.numpy(), k_vals=np.array(k_vals),
        recall_q_from_g=r_ab, recall_g_from_q=r_ba, top11_dist=D11, top11_idx=I11,
        rank0=O.rank0_exact(T, V), rank0_dot=O.rank0_exact(T, V, metric=O.METRIC_DOT),
        rank0_rev=O.rank0_exact(V, T),
    )
    print("retrieval_small:", r_ab, r_ba)


def gen_clip_loss(loss_mod):
    out = {}
    for name, (b, D, s) in {"c2_s100": (256, 512, 100.0), "c2_s14": (256, 512, 1.0 / 0.07),
                            "small": (32, 64, 20.0), "ragged": (200, 512, 100.0)}.items():
        vis, txt = make_batch_pair(b, D, seed=1023)
        sim = torch.tensor(s) * vis @ txt.t()               # model/model.py:369
        loss = loss_mod.clip_loss((vis, txt, sim), {})       # model/loss.py:18-22
        sim_g = sim.clone().requires_grad_(True)
        loss_mod.clip_loss((vis, txt, sim_g), {}).backward()
        out[name + "_loss"] = np.float32(loss.item())
        out[name + "_cfg"] = np.array([b, D, s], dtype=np.float64)
        out[name + "_sha"] = sha(vis.numpy(), txt.numpy())
        out[name + "_dsim_absmax"] = np.float32(sim_g.grad.abs().max().item())
        out[name + "_dsim_rowsum0"] = sim_g.grad[0].numpy().copy() if b <= 32 else sim_g.grad[0, :8].numpy().copy()
        if b <= 32:
            out[name + "_vis"] = vis.numpy()
            out[name + "_txt"] = txt.numpy()
            out[name + "_sim"] = sim.numpy()
            out[name + "_dsim"] = sim_g.grad.numpy()
        print("clip_loss", name, loss.item())
    np.savez_compressed(os.path.join(OUT, "clip_loss.npz"), **out)


def gen_cam(model_mod):
    out = {}
    torch.manual_seed(0)
    for name, (b, nc, D, layers, heads, rerand, avg) in {
        "c2_init": (256, 5, 512, 2, 8, False, True),       # reference zero-inits: closed form
        "c2_rand": (256, 5, 512, 2, 8, True, True),        # re-randomised: transformer is live
        "c2_linear": (64, 5, 512, 2, 8, True, False),      # final_linear(token 0) readout
        "small": (8, 3, 64, 2, 2, True, True),             # stored inputs + weights
    }.items():
        params = O.make_cam_params(D, layers, heads, seed=1023, rerandomise=rerand)
        g = torch.Generator().manual_seed(99)
        flw = torch.randn(D, D, generator=g) / D ** 0.5
        cam = RS.make_ref_cam(D, layers, heads, params=params, init_from_avg=avg,
                              final_linear_weight=flw)
        main, aux = make_cam_inputs(b, nc, D, seed=1023)
        with torch.no_grad():
            adapted = cam._adapt_feature(main, aux)                      # model/model.py:141-205
            tfm = cam.final_transformer(O.normalize(torch.stack([main, *aux], 0)))
        out[name + "_cfg"] = np.array([b, nc, D, layers, heads, int(rerand), int(avg)])
        out[name + "_sha"] = sha(main.numpy(), aux.numpy(), *[params[k].numpy() for k in sorted(params)])
        out[name + "_adapted"] = adapted.numpy() if b <= 64 else adapted[:32].numpy()
        out[name + "_tfm_tok0"] = tfm[0, :8].numpy()
        if name == "small":
            out["small_main"] = main.numpy()
            out["small_aux"] = aux.numpy()
            out["small_flw"] = flw.numpy()
            for k, v in params.items():
                out["small_param/" + k] = v.numpy()
        print("cam", name, adapted.shape, float(adapted.norm(dim=-1).mean()))
    # residual-activation variants (model/model.py:30-77) on the stored small case
    b, nc, D, layers, heads = 8, 3, 64, 2, 2
    params = O.make_cam_params(D, layers, heads, seed=1023, rerandomise=True)
    main, aux = make_cam_inputs(b, nc, D, seed=1023)
    g = torch.Generator().manual_seed(7)
    run_mean = 0.05 * torch.randn(D, generator=g)
    run_var = 0.5 + torch.rand(D, generator=g)
    out["act_running_mean"] = run_mean.numpy()
    out["act_running_var"] = run_var.numpy()
    for act in ("normalize", "squash", "squash10", "squash1p5", "tanh", "sub_mean", "bn"):
        cam = RS.make_ref_cam(D, layers, heads, params=params, residual_activation=act)
        cam._common_init()                                  # creates mean_center_bn when needed
        if act in ("sub_mean", "bn"):
            cam.mean_center_bn.running_mean.copy_(run_mean)
            cam.mean_center_bn.running_var.copy_(run_var)
        cam.eval()
        with torch.no_grad():
            out["act_" + act] = cam._adapt_feature(main, aux).numpy()
        print("cam act", act, float(np.abs(out["act_" + act]).max()))
    np.savez_compressed(os.path.join(OUT, "cam.npz"), **out)


def main():
    assert RS.reference_available(), "generate_golden.py needs /root/reference"
    loss_mod = RS.ref_loss_module()
    metric_mod = RS.ref_metric_module()
    model_mod = RS.ref_model_module()
    reval_mod = RS.ref_retrieval_evaluation_module()
    gen_retrieval_c1(metric_mod, reval_mod)
    gen_retrieval_small(metric_mod)
    gen_clip_loss(loss_mod)
    gen_cam(model_mod)


if __name__ == "__main__":
    main()
