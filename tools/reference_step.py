"""Dev tool: the reference's own Deformation / Engine.train, stock vs ptk_b200.install(), numbers printed
(tests/test_reference_gpu.py asserts on the same quantities).

    python tools/reference_step.py [B] [--layers N] [--hidden H] [--points P] [--cpu-dry]

--cpu-dry: no GPU -- `.cuda()` becomes a no-op and only the reference arm runs (checks the harness itself).
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def main():
    argv = sys.argv[1:]
    dry = "--cpu-dry" in argv
    opt = lambda name, default: int(argv[argv.index(name) + 1]) if name in argv else default
    B = int(argv[0]) if argv and argv[0].isdigit() else 2
    if dry:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        _rand = torch.rand
        torch.rand = lambda *a, **k: _rand(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    import ref_harness as H
    H.strict_fp32()
    ref = H.Reference()
    args = H.c3_args(num_GCN_layers=opt("--layers", 20), hidden_GCN_size=opt("--hidden", 300),
                     number_points=opt("--points", 3000))
    batch = H.make_batch(args, B=B, seed=0)
    t0 = time.time()
    info, mesh, net = ref.build(args, patched=False)
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    print(f"reference arm built in {time.time() - t0:.1f}s: adj {tuple(info['adj'].shape)} faces {tuple(info['faces'].shape)}")
    t0 = time.time()
    v_ref, l_ref, g_ref = ref.step(args, net, info, mesh, batch, patched=False)
    print(f"reference arm step {time.time() - t0:.1f}s loss {float(l_ref):.6f}")
    del net
    info64, mesh64, net64 = ref.build(args, patched=False, state=state, dtype=torch.float64)
    v_64, l_64, g_64 = ref.step(args, net64, info64, mesh64, batch, patched=False)
    print(f"reference fp64 loss {float(l_64):.8f}; fp32 reference vs fp64: verts {H.rel_err(v_ref, v_64):.2e} "
          f"loss {H.rel_err(l_ref, l_64):.2e}")
    del net64
    if dry:
        rows, zero = H.grad_report(g_ref, g_ref, g_64)
        print(f"reference fp32 grads vs fp64: worst {max(r[2] for r in rows):.2e}; zero-grad noise {max(z[1] for z in zero):.2e}")
        return
    info, mesh, net = ref.build(args, patched=True, state=state)
    t0 = time.time()
    v_b, l_b, g_b = ref.step(args, net, info, mesh, batch, patched=True)
    torch.cuda.synchronize()
    print(f"B200 arm step {time.time() - t0:.2f}s loss {float(l_b):.6f}")
    print(f"verts  direct {H.rel_err(v_b, v_ref):.2e}  vs fp64: b200 {H.rel_err(v_b, v_64):.2e} ref {H.rel_err(v_ref, v_64):.2e}")
    print(f"loss   direct {H.rel_err(l_b, l_ref):.2e}  vs fp64: b200 {H.rel_err(l_b, l_64):.2e} ref {H.rel_err(l_ref, l_64):.2e}")
    rows, zero = H.grad_report(g_b, g_ref, g_64)
    print("grads  (direct, b200 vs fp64, ref vs fp64), worst 12 of", len(rows), "+", len(zero), "structurally zero")
    for r in rows[:12]:
        print("   %.2e %.2e %.2e  %s" % r)
    print("grads  max direct %.2e, max b200-vs-fp64 %.2e, max ref-vs-fp64 %.2e" % tuple(max(r[i] for r in rows) for i in range(3)))
    print("grads  H1 violations:", [(k, d, eb, er) for d, eb, er, k in rows if not (d < 1e-5 or eb <= 2 * max(er, 1e-5))])
    print("zero-gradient noise / scale, worst: b200 %.2e ref %.2e" % (max(z[0] for z in zero), max(z[1] for z in zero)))
    del net
    # Engine.train, three Adam steps
    args2 = H.c3_args(num_GCN_layers=6, hidden_GCN_size=120, number_points=2000)
    batches = [H.make_batch(args2, B=2, seed=s) for s in (1, 2, 3)]
    info, mesh, net = ref.build(args2, patched=False)
    state0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    steps_ref, _, final_ref = ref.engine_train(args2, net, info, mesh, batches, patched=False)
    info64, mesh64, net64 = ref.build(args2, patched=False, state=state0, dtype=torch.float64)
    steps_64, _, final_64 = ref.engine_train(args2, net64, info64, mesh64, batches, patched=False)
    info, mesh, net = ref.build(args2, patched=True, state=state0)
    steps_b, _, final_b = ref.engine_train(args2, net, info, mesh, batches, patched=True)
    print("Engine.train losses per step\n  b200", steps_b, "\n  ref ", steps_ref, "\n  fp64", steps_64)
    worst = []
    lr_steps = args2.lr * 3
    for k, v64 in final_64.items():
        if not v64.is_floating_point() or "running" in k or "num_batches" in k:
            continue
        d64 = v64 - state0[k].double()
        if float(d64.abs().max()) < 0.1 * lr_steps:
            continue
        e_b = float((final_b[k].double() - v64).abs().max())
        e_r = float((final_ref[k].double() - v64).abs().max())
        worst.append((e_b / max(e_r, 1e-30), e_b, e_r, k))
    worst.sort(reverse=True)
    print("final weights: |b200 - fp64| / |ref - fp64|, worst 8 of", len(worst))
    for w in worst[:8]:
        print("   %.2f  %.2e %.2e  %s" % w)


if __name__ == "__main__":
    main()
