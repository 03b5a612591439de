"""Times the vocoder stage on the decoder bench batch (public 22 kHz / 80-band BigVGAN, random-init weights):
python scripts/vocoder_time.py [--batch 16] [--frames 512] [--iters 5]"""

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    from unitspeech_b200 import BigVGAN
    from unitspeech_b200.synthetic import PUBLIC_VOCODER_CONFIG, vocoder_state
    voc = BigVGAN(dict(PUBLIC_VOCODER_CONFIG))
    voc.load_state_dict(vocoder_state(PUBLIC_VOCODER_CONFIG, seed=1))
    voc.cuda().eval()
    voc.max_frames_per_call = a.batch * a.frames
    mel = (torch.randn(a.batch, 80, a.frames, device="cuda") * 2 - 4)
    for _ in range(2):
        out = voc(mel)
    torch.cuda.synchronize()
    l0 = voc.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        out = voc(mel)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    frames = a.batch * a.frames
    voc.set_profiling(True)
    voc(mel)
    torch.cuda.synchronize()
    prof = voc.get_profile()
    voc.set_profiling(False)
    cm, cw, cn = prof["conv_igemm"]
    am, aw, an = prof["snake_act"]
    om, ow, on = prof["other"]
    print(json.dumps({"conv_ms": cm, "conv_tflops_padded": cw / cm / 1e9, "conv_launches": cn, "act_ms": am,
                      "act_GBps": aw / am / 1e6, "act_launches": an, "other_ms": om, "other_GBps": ow / om / 1e6}))
    print(json.dumps({"vocoder_ms": ms, "mel_frames_per_s": frames / ms * 1e3, "audio_s_per_s": frames * 256 / 22050 / ms * 1e3,
                      "tensor_tflops": voc.flops_per_call / ms / 1e9, "launches_per_call": (voc.launch_count - l0) // a.iters,
                      "workspace_gb": voc.workspace_bytes / 2 ** 30, "finite": bool(torch.isfinite(out).all())}))


if __name__ == "__main__":
    main()
