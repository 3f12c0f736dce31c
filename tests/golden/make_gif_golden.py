#!/usr/bin/env python
"""Turns the reference's own rendered recordings into text fixtures (tests/golden/reference_gif_frames.json).

data/gif/*.gif are outputs of the reference's Rust code: act2gif (act2gif/src/{main,draw,term_image,font}.rs) builds the
game from a config, feeds it an action list and, for every Reaction::Redraw, renders the terminal (message line, dungeon,
status line; uilib/src/lib.rs process_reaction) with UbuntuMono into one GIF frame. All three recordings are the 32x16
mini dungeon of data/learned/ddqn-minidungeon/config.json (seed 5; image size = (W * (font + 1) / 2, H * (font + 1))):

  ddqn-small-16.gif / ddqn-small-24.gif   the first 30 actions of data/learned/ddqn-minidungeon/best-actions.json (the
                                          tool's default --max 30) at font size 16 / 24: 28 frames
  pporesnet-cog19-10seed.gif              another agent on the same game, 48 frames; its action list is not in the
                                          repository

This script reads the frames back: a character cell is the (font/2) x font pixel box at (col * font / 2,
row * font + font / 4) (font.rs draw_range), every cell is classified against the same TTF rendered with PIL (normalised
correlation; '~' is left out of the alphabet: at 16 px it is indistinguishable from '-' and the game never prints it).
Checks made here: the 16 px and 24 px recordings decode to the same text; for the second recording the action sequence
is inferred with the oracle - at every frame EXACTLY ONE of the ten single actions (8 moves, search, '>') reproduces the
dungeon rows; a '>' off the stairs draws nothing and is read off the message line - and stored with the frames.

Needs /root/reference, PIL and the built oracle; the tests only read the JSON it writes.
"""
import json
import os
import sys

import numpy as np
from PIL import Image, ImageDraw, ImageFont

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
TTF = os.path.join(REF, "data/fonts/UbuntuMono-R.ttf")
BG = np.array([0, 43, 54], float)       # act2gif/src/theme.rs SR_BASE03 (solarized-dark, the default theme)
FG = np.array([147, 161, 161], float)   # SR_BASE1
W, H = 32, 16


def templates(scale):
    font = ImageFont.truetype(TTF, scale)
    w, h, pad = scale // 2, scale, scale // 4
    out = {}
    for code in range(32, 127):
        if chr(code) == "~":
            continue
        img = Image.new("L", (w * 3, 2 * h + pad), 0)
        ImageDraw.Draw(img).text((w, h), chr(code), font=font, fill=255, anchor="ls")  # baseline at y = font (font.rs:46)
        out[chr(code)] = (np.array(img).astype(float) / 255.0)[pad:pad + h, w:2 * w]
    return out


def decode(path, scale):
    T = templates(scale)
    chars = list(T)
    tm = np.stack([T[c].ravel() for c in chars])
    tn = np.linalg.norm(tm, axis=1)
    w, h, pad = scale // 2, scale, scale // 4
    im = Image.open(path)
    assert im.size == (W * (scale + 1) // 2, H * (scale + 1)), im.size  # term_image.rs:33-34
    d = FG - BG
    frames, worst = [], 1.0
    for i in range(im.n_frames):
        im.seek(i)
        a = np.clip(((np.array(im.convert("RGB")).astype(float) - BG) @ d) / (d @ d), 0, 1)  # coverage of the font colour
        rows = []
        for r in range(H):
            line = ""
            for c in range(W):
                v = a[r * h + pad:r * h + pad + h, c * w:(c + 1) * w].ravel()
                if v.max() < 0.25:
                    line += " "
                    continue
                sc = (tm @ v) / (tn * np.linalg.norm(v) + 1e-9)
                k = int(sc.argmax())
                worst = min(worst, float(sc[k]))
                line += chars[k]
            rows.append(line)
        frames.append(rows)
    return frames, worst


def infer_actions(cfg, frames):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    oracle_py.build()
    keys = ""
    no_stair = "Hmm... there seems to be no downstair"[:W]
    for i, fr in enumerate(frames):
        if fr[0] == no_stair and (i == 0 or frames[i - 1][0] != no_stair):
            keys += ">"  # a '>' off the stairs: Notify only, no Redraw, no frame of its own - the message line shows it
        ok = []
        for k in "hjklyubns>":
            o = oracle_py.OracleEnv(cfg, max_steps=2000)
            for q in keys + k:
                o.react(ord(q))
            if o.dungeon()[1:H - 1] == fr[1:H - 1]:
                ok.append(k)
        assert len(ok) == 1, "frame %d: %d single actions reproduce it (%s)" % (i, len(ok), ok)
        keys += ok[0]
    return keys


def main():
    with open(os.path.join(REF, "data/learned/ddqn-minidungeon/config.json")) as f:
        cfg = json.load(f)
    gif = os.path.join(REF, "data/gif")
    f16, s16 = decode(os.path.join(gif, "ddqn-small-16.gif"), 16)
    f24, s24 = decode(os.path.join(gif, "ddqn-small-24.gif"), 24)
    assert f16 == f24, "the two recordings of the same run decode differently"
    ppo, sp = decode(os.path.join(gif, "pporesnet-cog19-10seed.gif"), 16)
    ppo_keys = infer_actions(cfg, ppo)
    out = {
        "how": "tests/golden/make_gif_golden.py (glyph-by-glyph decode of the reference's data/gif/*.gif, see its docstring)",
        "config": cfg,
        "ddqn_small": {
            "cite": "data/gif/ddqn-small-{16,24}.gif = act2gif of the first 30 actions of data/learned/ddqn-minidungeon/"
                    "best-actions.json (act2gif/src/main.rs:66-73 default --max 30; one frame per Reaction::Redraw, draw.rs:56-66)",
            "n_actions": 30, "worst_glyph_score": [s16, s24], "frames": f24,
        },
        "ppo_cog19": {
            "cite": "data/gif/pporesnet-cog19-10seed.gif: same config, actions not shipped; `keys` inferred with the oracle, "
                    "exactly one candidate per frame",
            "keys": ppo_keys, "worst_glyph_score": [sp], "frames": ppo,
        },
    }
    dst = os.path.join(HERE, "reference_gif_frames.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote %s: %d + %d frames, glyph scores >= %.2f, inferred keys %s" % (dst, len(f24), len(ppo), min(s16, s24, sp), ppo_keys))


if __name__ == "__main__":
    main()
