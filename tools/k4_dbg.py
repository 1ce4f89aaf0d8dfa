"""Where does the strip kernel differ from the oracle?  4K frames into a 16384^2 canvas (many chunks per CTA), one update at a time."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mosaicking_b200 as mb
from oracle import oracle as orc

code = 1
rng = np.random.default_rng(100 + code)
Hc = Wc = 16384
mp = mb.Mapper(Wc, Hc, alpha_blend=0.5, interp="linear")
Hs = []
for k in range(3):
    th = 0.11 * k - 0.05
    Hs.append(np.array([[np.cos(th), -np.sin(th), 9000.3 + 900 * k], [np.sin(th), np.cos(th), 11000.7 + 500 * k], [1e-6 * k, 5e-7 * k, 1.0]]))


def _img(h, w, black):
    img = rng.integers(1, 256, (h, w, 3), dtype=np.uint8)
    img[rng.random((h, w)) < black] = 0
    return img


imgs = [_img(2160, 3840, 0.01) for _ in range(3)]
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    mp2 = mb.Mapper(Wc, Hc, alpha_blend=0.5, interp="linear")
    regions = []
    win = wm = None
    for k, (img, H) in enumerate(zip(imgs, Hs)):
        mp2.update(img, H)
        torch.cuda.synchronize()
        regions.append(mp2.last_region)
        x0 = min(r[0] for r in regions); y0 = min(r[1] for r in regions); x1 = max(r[2] for r in regions); y1 = max(r[3] for r in regions)
        x0, y0, x1, y1 = max(0, x0 - 16), max(0, y0 - 16), min(Wc, x1 + 16), min(Hc, y1 + 16)
        win = np.zeros((y1 - y0, x1 - x0, 3), np.uint8); wm = np.zeros(win.shape[:2], np.uint8)
        for img2, H2 in zip(imgs[:k + 1], Hs[:k + 1]):
            orc.mapper_update_window(win, wm, (Hc, Wc), (x0, y0), img2, H2, 0.5, code)
        out = mp2.output_device[y0:y1, x0:x1].cpu().numpy(); msk = mp2._canvas_mask[y0:y1, x0:x1].cpu().numpy()
        d = np.argwhere((out != win).any(axis=2)); dm = np.argwhere(msk != wm)
        print(f"rep {rep} after update {k}: {len(d)} px differ, {len(dm)} mask; window origin ({x0},{y0})")
        if len(d):
            ys, xs = d[:, 0] + y0, d[:, 1] + x0
            print("   x range", xs.min(), xs.max(), "y range", ys.min(), ys.max())
            strips = np.unique(xs // 64)
            print("   strips", strips[:20].tolist(), "n", len(strips))
            for s in strips[:6]:
                sel = xs // 64 == s
                yy = np.unique(ys[sel])
                print(f"   strip {s}: rows {yy[:12].tolist()} .. {yy[-3:].tolist()} ({len(yy)} rows), cols {np.unique(xs[sel] % 64)[:10].tolist()}")
            prev = np.zeros_like(win); pm = np.zeros_like(wm)
            for img2, H2 in zip(imgs[:k], Hs[:k]):
                orc.mapper_update_window(prev, pm, (Hc, Wc), (x0, y0), img2, H2, 0.5, code)
            only = np.zeros_like(win); om = np.zeros_like(wm)
            orc.mapper_update_window(only, om, (Hc, Wc), (x0, y0), imgs[k], Hs[k], 0.0, code)     # alpha 0: the warped frame itself
            cls = {"eq_prev": 0, "eq_warped": 0, "other": 0}
            for (yy, xx) in d:
                g = out[yy, xx]
                if (g == prev[yy, xx]).all(): cls["eq_prev"] += 1
                elif (g == only[yy, xx]).all(): cls["eq_warped"] += 1
                else: cls["other"] += 1
            print("   classes", cls)
            for (yy, xx) in d[:8]:
                print("   ", (int(xx + x0), int(yy + y0)), "gpu", out[yy, xx].tolist(), "want", win[yy, xx].tolist(), "prev", prev[yy, xx].tolist(), "warped", only[yy, xx].tolist(), "pm", int(pm[yy, xx]))
            break
    mp2.release()
