"""Tuning aid: times conv shapes of the bench workload (Be=48, T=512) under USB_DBG_* overrides."""
import ctypes, os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

SHAPES = {  # name: kind, N, H, W, C0, C1, Cout, stats
    "l0_128": (0, 48, 80, 512, 128, 0, 128, 1),
    "l1_256": (0, 48, 40, 256, 256, 0, 256, 1),
    "l1_128to256": (0, 48, 40, 256, 128, 0, 256, 1),
    "l2_512": (0, 48, 20, 128, 512, 0, 512, 1),
    "l3_1024": (0, 48, 10, 64, 1024, 0, 1024, 1),
    "u2_512to128": (0, 48, 40, 256, 256, 256, 128, 1),
    "u3_256to128": (0, 48, 80, 512, 128, 128, 128, 1),
    "down_l0": (1, 48, 80, 512, 128, 0, 128, 0),
    "down_l1": (1, 48, 40, 256, 256, 0, 256, 0),
    "up_l2": (3, 48, 20, 128, 256, 0, 256, 0),
    "l3_512": (0, 48, 10, 64, 512, 0, 512, 1),
    "qkv_l0": (2, 48, 80, 512, 128, 0, 384, 0),
    "up_l1": (3, 48, 40, 256, 128, 0, 128, 0),
    "qkv_l1": (2, 48, 40, 256, 256, 0, 384, 0),
    "res_l1": (2, 48, 40, 256, 128, 0, 256, 0),
    "res_u2": (2, 48, 40, 256, 256, 256, 128, 0),
    # the per-GPU shard of BASELINE configs[2] (Be = 96, T = 1000): strided / transposed convs
    "up_l1_T1000": (3, 96, 40, 500, 128, 0, 128, 0),
    "up_l2_T1000": (3, 96, 20, 250, 256, 0, 256, 0),
    "up_l3_T1000": (3, 96, 10, 125, 512, 0, 512, 0),
    "down_l0_T1000": (1, 96, 80, 1000, 128, 0, 128, 0),
    "down_l1_T1000": (1, 96, 40, 500, 256, 0, 256, 0),
    "l0_128_T1000": (0, 96, 80, 1000, 128, 0, 128, 1),
}

def run_one(names):
    from gpu_ops import OpHandle
    h = OpHandle(0)
    out = {}
    for n in names:
        k, N, H, W, C0, C1, Co, st = SHAPES[n]
        ms = ctypes.c_float()
        from unitspeech_b200 import abi
        abi.check(h.lib.usb_dbg_conv_time(h.h, k, N, H, W, C0, C1, Co, st, 5, ctypes.byref(ms)))
        taps = {0: 9, 1: 9, 2: 1, 3: 16}[k]
        px = N * H * W * (0.25 if k == 1 else 1)
        fl = 2 * px * Co * (C0 + C1) * taps
        out[n] = (round(ms.value * 1e3, 1), round(fl / ms.value / 1e9, 1))
    print(json.dumps(out))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        run_one(sys.argv[2:])
    else:
        configs = [{"USB_NO_HALO": "1"}, {}, {"USB_DBG_STAGES": "6"}, {"USB_DBG_STAGES": "5"}, {"USB_DBG_FLAGS": "2"},
                   {"USB_NO_HALO": "1", "USB_DBG_FLAGS": "2"}] if os.environ.get("SWEEP") == "halo" else [{}, {"USB_DBG_FLAGS": "2"}] if os.environ.get("SWEEP") == "epi" else [{}, {"USB_NO_SWAP_AB": "1"}] if os.environ.get("SWEEP") == "swap" else [{}, {"USB_DBG_FLAGS": "1"}, {"USB_DBG_FLAGS": "2"}, {"USB_DBG_FLAGS": "3"}, {"USB_DBG_FLAGS": "4"},
                   {"USB_DBG_FLAGS": "6"}, {"USB_DBG_FLAGS": "7"},
                   {"USB_DBG_STAGES": "3"}, {"USB_DBG_STAGES": "4"}, {"USB_DBG_BH": "2"}, {"USB_DBG_BH": "4"},
                   {"USB_DBG_BH": "8"}, {"USB_DBG_BH": "1"}]
        names = sys.argv[1:] or list(SHAPES)
        for c in configs:
            env = dict(os.environ); env.update(c)
            r = subprocess.run([sys.executable, __file__, "child"] + names, env=env, capture_output=True, text=True)
            print(c, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
