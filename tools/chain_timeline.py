"""Per-step phase timeline of the decode chain kernel (b200q_debug_set_chain_timeline): Llama-2-7B shapes, `blocks`
decoder blocks in one launch.  Prints [min, median, max] over CTAs in us relative to the kernel start."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qllm_b200
import bench

def main(blocks=6, reps=3):
    dev = torch.device("cuda:0")
    bench.BLOCKS = blocks
    model = bench.build_model(dev, 0, 1, "GEMM")
    step = bench.ChainDecodeStep(model, dev, 1, blocks)
    step.h.copy_(torch.randn(1, bench.HIDDEN).half())
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(reps):
        step.run(s)
    torch.cuda.synchronize()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n_steps = 4 * blocks
    buf = torch.zeros(n_steps * sms * 16, dtype=torch.int64, device=dev)
    qllm_b200.lib.b200q_debug_set_chain_timeline(buf.data_ptr())
    step.run(s)
    torch.cuda.synchronize()
    qllm_b200.lib.b200q_debug_set_chain_timeline(None)
    t = buf.cpu().numpy().reshape(n_steps, sms, 16).astype(np.float64)
    t0 = t[0, :, 0][t[0, :, 0] > 0].min()
    names = ["start", "x_ready", "digits", "own_units", "all_units", "stored", "fin_start", "y_written", "prod_first", "prod_last"]
    kinds = ["qkv", "o", "gate|up", "down"]
    for g in range(n_steps):
        row = f"{g:3d} {kinds[g % 4]:8s}"
        for i, nm in enumerate(names):
            v = t[g, :, i]
            v = v[v > 0]
            if len(v) == 0:
                continue
            v = (v - t0) / 1e3
            row += f" {nm}[{v.min():7.2f},{np.median(v):7.2f},{v.max():7.2f}]"
        st = t[g, :, 10] / 1e3
        row += f" stall_us[med {np.median(st):5.2f} max {st.max():5.2f}]"
        nu = np.maximum(t[g, :, 12], 1)
        row += f" w0: units {np.median(t[g, :, 12]):.0f} wait_cyc/unit {np.median(t[g, :, 11] / nu):.0f} loop_cyc/unit {np.median(t[g, :, 13] / nu):.0f}"
        print(row)
    total = (t[n_steps - 1, :, 7].max() - t0) / 1e3
    print(f"total {total:.2f} us for {blocks} blocks = {total / blocks:.2f} us / block")

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 6)
