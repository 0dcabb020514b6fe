"""CUDA-graph replay of the config-5 bPC step (amortiser sweep + T fused inference steps + parameter gradients, all through
the functional API on fixed tensors) vs plain stream launches."""
import sys, torch
sys.path.insert(0, '.')
import bench, jpc_b200 as J
w = bench.WORKLOADS['c5']
dev = torch.device('cuda')
gen_l, _, _ = bench.make_inputs(dict(w, d_in=w["d_in"], d_out=w["d_out"]), seed_w=0)
amo_l, _, _ = bench.make_inputs(dict(w, d_in=w["d_out"], d_out=w["d_in"]), seed_w=1)
g = torch.Generator().manual_seed(1)
img = torch.randn(w["B"], w["d_out"], generator=g).to(dev)
lab = torch.nn.functional.one_hot(torch.randint(0, w["d_in"], (w["B"],), generator=g), w["d_in"]).float().to(dev)
td, amort = bench.to_device_model(gen_l, dev), bench.to_device_model(amo_l, dev)
bu = amort[::-1]
T = bench.n_steps(w)
for dtype in ("bf16", "f32x3"):
    J.set_compute_dtype(dtype)
    def step():
        acts = J.init_activities_with_ffwd(amort, img)
        acts = J.solve_bpc_inference(td, bu, acts, img, input=lab, n_steps=T, dt=w["dt"])
        return J.compute_bpc_param_grads(td, bu, acts, img, x=lab)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3): out = step()
        side.synchronize()
        ref = out[0][0][1].weight.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): step()
        e1.record(); side.synchronize()
        plain = e0.elapsed_time(e1) / 20
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            out = step()
        gr.replay(); side.synchronize()
        ok = torch.equal(out[0][0][1].weight, ref)
        e0.record()
        for _ in range(20): gr.replay()
        e1.record(); side.synchronize()
        print(f"c5 {dtype}: stream launches {plain:.3f} ms/step, graph replay {e0.elapsed_time(e1) / 20:.3f} ms/step, same result: {ok}")
