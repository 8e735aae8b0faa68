import sys, os, time, torch
sys.path.insert(0, os.getcwd())
from bench import REF_ARCH
from audio_diffusion_b200.schedulers import DDPMScheduler
from audio_diffusion_b200.unet import UNet2DModel
dev = torch.device("cuda:0")
for res in (256, 64):
    model = UNet2DModel(sample_size=(res, res), seed=0, **REF_ARCH).to(dev)
    sch = DDPMScheduler(); sch.set_timesteps(1000)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(1, 1, res, res, generator=g, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        def eager(i):
            t = sch.timesteps[i]; z = torch.randn(x.shape, generator=g, device=dev)
            model.forward_step(x, t, sch.step_coef(t), noise=z, out=x)
        for i in range(5): eager(i)
        torch.cuda.synchronize(); e0.record()
        for i in range(30): eager(5 + i)
        e1.record(); torch.cuda.synchronize(); te = e0.elapsed_time(e1) / 30
        st = model.graph_stepper(x)
        def graph(i):
            t = sch.timesteps[i]; z = torch.randn(x.shape, generator=g, device=dev)
            st.step(t, sch.step_coef(t), z)
        for i in range(5): graph(40 + i)
        torch.cuda.synchronize(); e0.record()
        for i in range(30): graph(45 + i)
        e1.record(); torch.cuda.synchronize(); tg = e0.elapsed_time(e1) / 30
    print(f"batch 1 {res}x{res}: eager {te:.3f} ms/step, graph {tg:.3f} ms/step")
