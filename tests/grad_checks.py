"""TEST INFRASTRUCTURE shared by tests/test_autograd_cpu.py (oracle-backed stand-in for the CUDA entry points) and
tests/test_train_gpu.py (the real kernels): run a module forward + backward on a golden case and compare every gradient
with what autograd through the UNMODIFIED reference produced (oracle/make_golden.py, loss = Σ out ⊙ cotangent(seed))."""
import numpy as np
import torch

from oracle import cases, oracle_torch


def tsd(sd, dev="cpu"):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in sd.items()}


def coo(mats, dev="cpu"):
    return [oracle_torch.to_torch_coo(m).to(dev) for m in mats]


def loss_of(outs, seed):
    return sum((o * torch.from_numpy(cases.cotangent(seed + j, tuple(o.shape))).to(o.device)).sum() for j, o in enumerate(outs))


def stack3(v):
    return torch.stack(list(v)) if isinstance(v, (list, tuple)) else (v if v.dim() == 3 else v[None])


def check_grads(mod, case, tol, extra=None):
    want = dict(case["grads"])
    assert want, "case stores no reference gradients"
    got = {name: p.grad for name, p in mod.named_parameters() if p.grad is not None}
    if extra:
        got.update(extra)
    assert sorted(got) == sorted(want), (sorted(got), sorted(want))
    worst = 0.0
    for name, g in got.items():
        err = cases.relerr(g.detach().cpu().numpy(), want[name])
        assert np.isfinite(err) and err < tol, (name, err)
        worst = max(worst, err)
    return worst


def core_diffusion_grad(pkg, c, dev, tol):
    m = c["meta"]
    mod = pkg.CoreDiffusion(m["d_in"], m["d_out"], bias=m["bias"], rnn_type=m["rnn_type"]).to(dev)
    mod.load_state_dict(tsd(c["sd"], dev), strict=True)
    x = torch.from_numpy(c["x"]).to(dev).requires_grad_(True)
    y = mod(x, coo(c["adj"], dev))
    loss_of([y], m["cot_seed"]).backward()
    assert mod.linear.weight.grad is None                      # unused parameter (layers.py:46), like the reference
    return check_grads(mod, c, tol, {"x": x.grad})


def mlp_grad(pkg, c, dev, tol):
    m = c["meta"]
    mod = pkg.MLP(m["d_in"], m["hid"], m["d_out"], m["layer_num"], bias=m["bias"], activate_type=m["act"]).to(dev)
    mod.load_state_dict(tsd(c["sd"], dev), strict=True)
    if isinstance(c["x"], np.ndarray):
        x = torch.from_numpy(c["x"]).to(dev).requires_grad_(True)
        loss_of([mod(x)], m["cot_seed"]).backward()
        return check_grads(mod, c, tol, {"x": x.grad})
    loss_of([mod(oracle_torch.to_torch_coo(c["x"]).to(dev))], m["cot_seed"]).backward()
    return check_grads(mod, c, tol)


def cdn_grad(pkg, c, dev, tol):
    m = c["meta"]
    mod = pkg.CDN(m["d_in"], m["hid"], m["d_out"], m["diffusion_num"], rnn_type=m["rnn_type"]).to(dev)
    mod.load_state_dict(tsd(c["sd"], dev), strict=True)
    x = torch.from_numpy(c["x"]).to(dev).requires_grad_(True)
    loss_of([mod(x, coo(c["adj"], dev))], m["cot_seed"]).backward()
    return check_grads(mod, c, tol, {"x": x.grad})


def build_model(pkg, m, dev):
    if m["kind"] == "ctgcn":
        mod = pkg.CTGCN(m["d_in"], m["hid"], m["d_out"], m["trans_num"], m["diffusion_num"], m["T"], rnn_type=m["rnn_type"],
                        model_type=m["model_type"], trans_activate_type=m["act"])
    else:
        mod = pkg.CGCN(m["d_in"], m["hid"], m["d_out"], m["trans_num"], m["diffusion_num"], rnn_type=m["rnn_type"],
                       model_type=m["model_type"], trans_activate_type=m["act"])
    return mod.to(dev)


def model_inputs(c, dev):
    xs = [torch.from_numpy(x).to(dev) if isinstance(x, np.ndarray) else oracle_torch.to_torch_coo(x).to(dev) for x in c["x_list"]]
    adj = [coo(al, dev) for al in c["adj_lists"]]
    return xs, adj


def model_grad(pkg, c, dev, tol, fwd_tol=0.0):
    m = c["meta"]
    mod = build_model(pkg, m, dev)
    mod.load_state_dict(tsd(c["sd"], dev), strict=True)
    xs, adj = model_inputs(c, dev)
    res = mod(xs[0], adj[0]) if m.get("single") else mod(xs, adj)
    out, trans = res if m["model_type"] == "S" else (res, None)
    loss_of([stack3(out)] + ([stack3(trans)] if trans is not None else []), m["cot_seed"]).backward()
    worst = check_grads(mod, c, tol)
    # the no-grad fast path (outputs written straight into the [N, T, D] buffer) gives the same forward values
    with torch.no_grad():
        res2 = mod(xs[0], adj[0]) if m.get("single") else mod(xs, adj)
    out2 = res2[0] if m["model_type"] == "S" else res2
    assert torch.allclose(stack3(out2), stack3(out).detach(), rtol=0, atol=fwd_tol)
    return worst
