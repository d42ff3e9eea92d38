"""Dev harness (GPU): CTC_Model forward/backward vs the CPU oracle model, plus first timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from ctc_pytorch_b200.model import CTC_Model
from ctc_pytorch_b200.loss import CTCLoss
from oracle.model_ref import RefAcousticModel, synthetic_batch

dev = "cuda"

def relerr(a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

def run(T, N, F, H, L, C, bn=True, S=10, seed=1, batch_tile=0, train=True, check_ref=True):
    torch.manual_seed(seed)
    rnn_param = {"rnn_input_size": F, "rnn_hidden_size": H, "rnn_layers": L, "rnn_type": nn.LSTM,
                 "bidirectional": True, "batch_norm": bn}
    m = CTC_Model(rnn_param=rnn_param, num_class=C, drop_out=0.0)
    ref = RefAcousticModel(F, H, L, C, batch_norm=bn)
    ref.load_state_dict(m.state_dict())
    m = m.to(dev); m.batch_tile = batch_tile
    x, frac, tg, tl = synthetic_batch(T, N, F, C, S, seed)
    il = (frac * T).long()
    tag = "T=%d N=%d F=%d H=%d L=%d C=%d bn=%d tile=%d" % (T, N, F, H, L, C, bn, batch_tile)
    try:
        # eval forward
        m.eval(); ref.eval()
        with torch.no_grad():
            o = m(x.to(dev)); torch.cuda.synchronize()
            if check_ref:
                r = ref(x)
                print("[%s] eval fwd maxabs=%.3e rel=%.3e" % (tag, (o.cpu() - r).abs().max().item(), relerr(o, r)))
        if not train: return
        m.train(); ref.train()
        lossf = CTCLoss(reduction="sum")
        o = m(x.to(dev))
        loss = lossf(o, tg.to(dev), il.to(dev), tl.to(dev)) / N
        loss.backward(); torch.cuda.synchronize()
        if check_ref:
            r = ref(x)
            lr = nn.CTCLoss(reduction="sum")(r, tg, il, tl) / N
            lr.backward()
            print("[%s] train fwd maxabs=%.3e loss=%.6f ref=%.6f rel=%.2e" % (tag, (o.detach().cpu() - r.detach()).abs().max().item(),
                  loss.item(), lr.item(), abs(loss.item() - lr.item()) / abs(lr.item())))
            rp = dict(ref.named_parameters())
            worst = 0
            for name, p in m.named_parameters():
                if p.grad is None: print("   %s: NO GRAD" % name); continue
                e = relerr(p.grad, rp[name].grad); worst = max(worst, e)
                print("   %-40s rel=%.3e |g|=%.3e" % (name, e, rp[name].grad.norm().item()))
            rb = dict(ref.named_buffers())
            for name, b in m.named_buffers():
                if "running" in name:
                    print("   %-40s rel=%.3e" % (name, relerr(b, rb[name])))
            print("[%s] worst grad rel err %.3e" % (tag, worst))
    except Exception as e:
        import traceback; print(traceback.format_exc()[-600:])
        print("[%s] EXCEPTION %s" % (tag, e))

def timeit(T, N, F, H, L, C, batch_tile=0, iters=3):
    torch.manual_seed(0)
    rnn_param = {"rnn_input_size": F, "rnn_hidden_size": H, "rnn_layers": L, "rnn_type": nn.LSTM,
                 "bidirectional": True, "batch_norm": True}
    m = CTC_Model(rnn_param=rnn_param, num_class=C, drop_out=0.0).to(dev); m.batch_tile = batch_tile
    x, frac, tg, tl = synthetic_batch(T, N, F, C, 60, 0)
    xd, tgd, ild, tld = x.to(dev), tg.to(dev), (frac * T).long().to(dev), tl.to(dev)
    lossf = CTCLoss(reduction="sum"); m.train()
    def step():
        m.zero_grad(set_to_none=True)
        o = m(xd); loss = lossf(o, tgd, ild, tld) / N; loss.backward(); return loss
    try:
        step(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(); o = m(xd); e1.record(); loss = lossf(o, tgd, ild, tld) / N; loss.backward(); e2.record(); torch.cuda.synchronize()
        print("TIME T=%d N=%d H=%d L=%d tile=%d: fwd %.2f ms, loss+bwd %.2f ms" % (T, N, H, L, batch_tile, e0.elapsed_time(e1), e1.elapsed_time(e2)))
        e0.record()
        for _ in range(iters): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print("TIME T=%d N=%d H=%d L=%d tile=%d: step %.2f ms -> %.1f utt/s" % (T, N, H, L, batch_tile, ms, N / ms * 1e3))
    except Exception as e:
        import traceback; print(traceback.format_exc()[-600:])

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    run(6, 3, 40, 128, 1, 10, bn=False, S=2)
    run(12, 3, 40, 128, 2, 10, bn=True, S=4)
    run(12, 20, 40, 256, 2, 10, bn=True, S=4)
    run(12, 20, 40, 256, 2, 10, bn=True, S=4, batch_tile=32)
    run(50, 32, 40, 512, 4, 62, bn=True, S=10)
    run(50, 32, 40, 512, 4, 62, bn=True, S=10, batch_tile=32)
    if not quick:
        run(40, 64, 40, 640, 2, 48, bn=True, S=8)
        for bt in (16, 32):
            timeit(800, 32, 40, 512, 4, 62, batch_tile=bt)
        timeit(1200, 64, 40, 640, 5, 48, batch_tile=0, iters=2)
