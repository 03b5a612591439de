"""BASELINE.json configs[3]: decoder (50-step CFG) + BigVGAN vocoder end to end, random-init weights, synthetic
conditioning.  python scripts/pipeline_time.py [--batch 64] [--frames 1000]"""

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    from unitspeech_b200 import BigVGAN, UnitSpeech, denormalize_mel
    from unitspeech_b200.synthetic import PUBLIC_VOCODER_CONFIG, random_init_state_dict, synthetic_inputs, vocoder_state
    dev = torch.device("cuda", 0)
    dec = UnitSpeech(80, 128, (1, 2, 4, 8), beta_min=0.05, beta_max=20, pe_scale=1000, spk_emb_dim=256)
    dec.load_state_dict(random_init_state_dict(dec))
    dec = dec.to(dev).eval()
    voc = BigVGAN(dict(PUBLIC_VOCODER_CONFIG))
    voc.load_state_dict(vocoder_state(PUBLIC_VOCODER_CONFIG, seed=1))
    voc.to(dev).eval()
    B, T, n = a.batch, a.frames, a.steps
    z, mask, cond, spk, noise = (t.to(dev) for t in synthetic_inputs(B, T, n))
    mel_min, mel_max = torch.full((80, 1), -11.5, device=dev), torch.full((80, 1), 2.0, device=dev)

    def run():
        y = dec(z, mask, cond, spk, n, text_gradient_scale=1.0, spk_gradient_scale=1.0, noise=noise)
        mel = denormalize_mel(y, mel_min, mel_max)               # inference.py:140
        return y, voc(mel).squeeze(1).clamp(-1, 1)               # inference.py:141

    # one untimed pass (plans, workspaces, lazy module load), then the timed one
    run()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    y = dec(z, mask, cond, spk, n, text_gradient_scale=1.0, spk_gradient_scale=1.0, noise=noise)
    e[1].record()
    audio = voc(denormalize_mel(y, mel_min, mel_max)).squeeze(1).clamp(-1, 1)
    e[2].record()
    torch.cuda.synchronize()
    d_ms, v_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    frames = B * T
    print(json.dumps({"workload": f"{B} utterances x {T} frames, {n}-step CFG decoder + BigVGAN (22 kHz, hop 256)",
                      "decoder_ms": d_ms, "vocoder_ms": v_ms, "pipeline_mel_frames_per_s": frames / (d_ms + v_ms) * 1e3,
                      "decoder_frames_per_s": frames / d_ms * 1e3, "vocoder_frames_per_s": frames / v_ms * 1e3,
                      "rtf": (d_ms + v_ms) / 1e3 / (frames * 256 / 22050.0), "audio_shape": list(audio.shape),
                      "finite": bool(torch.isfinite(audio).all())}))


if __name__ == "__main__":
    main()
