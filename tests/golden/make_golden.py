#!/usr/bin/env python3
"""Generate the committed golden fixtures from the reference's own artefacts.

Runs ONLY in the build container (needs /root/reference, which does not exist
on the GPU box).  Outputs (small, committed):

  g1_food_list.json        the 50-entry food_list + post-draw Xoshiro state stored in
                           /root/reference/trainers/very_long_training1.bson  (tr.game)
                           -> pins structs.jl:70 (food_list from Xoshiro(42))
  g2_boards_double3.npy    240 decoded 10x10 boards of
                           /root/reference/trainer_gifs/very_long_double_training3.gif
                           -> pins step!/sample_food!/update_board! (utils.jl:13-109)
  g5_boards_training1.npy  130 decoded boards of trainer_gifs/very_long_training1.gif
                           (older one-frame code; only its food cells are used)
  g4_bson_game.json        board / snake / score / lost stored in the BSON game object

Nothing here is reference SOURCE; these are data artefacts decoded to integers.
"""
import json
import os
import struct
import sys

import numpy as np
from PIL import Image

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------- BSON
def _cstring(b, o):
    e = b.index(b"\x00", o)
    return b[o:e].decode("utf8"), e + 1


def parse_doc(b, o=0, as_list=False):
    (n,) = struct.unpack_from("<i", b, o)
    end = o + n
    o += 4
    out = [] if as_list else {}
    while b[o] != 0:
        t = b[o]
        o += 1
        k, o = _cstring(b, o)
        if t == 0x01:
            (v,) = struct.unpack_from("<d", b, o); o += 8
        elif t == 0x02:
            (l,) = struct.unpack_from("<i", b, o); o += 4
            v = b[o:o + l - 1].decode("utf8"); o += l
        elif t == 0x03:
            v, o = parse_doc(b, o)
        elif t == 0x04:
            v, o = parse_doc(b, o, as_list=True)
        elif t == 0x05:
            (l,) = struct.unpack_from("<i", b, o); o += 5
            v = bytes(b[o:o + l]); o += l
        elif t == 0x08:
            v = bool(b[o]); o += 1
        elif t == 0x0A:
            v = None
        elif t == 0x10:
            (v,) = struct.unpack_from("<i", b, o); o += 4
        elif t == 0x12:
            (v,) = struct.unpack_from("<q", b, o); o += 8
        else:
            raise ValueError("bson type 0x%02x" % t)
        if as_list:
            out.append(v)
        else:
            out[k] = v
    assert o + 1 == end
    return out, end


def type_name(t, backrefs):
    while isinstance(t, dict) and t.get("tag") == "backref":
        t = backrefs[t["ref"] - 1]
    if isinstance(t, dict) and t.get("tag") == "datatype":
        return ".".join(t["name"])
    return None


def golden_from_bson():
    with open(os.path.join(REF, "trainers/very_long_training1.bson"), "rb") as f:
        raw = f.read()
    doc, _ = parse_doc(raw)
    backrefs = doc["_backrefs"]
    tr = doc["tr"]

    def deref(x):
        while isinstance(x, dict) and x.get("tag") == "backref":
            x = backrefs[x["ref"] - 1]
        return x

    tr = deref(tr)
    assert type_name(tr["type"], backrefs).endswith("Trainer")
    game = deref(tr["data"][0])
    assert type_name(game["type"], backrefs).endswith("SnakeGame")
    fields = [deref(x) for x in game["data"]]
    # locate arrays by element type
    food = rng = board = snake = None
    for i, f in enumerate(fields):
        if not isinstance(f, dict):
            continue
        tn = type_name(f.get("type"), backrefs) if "type" in f else None
        if f.get("tag") == "array" and tn and "CartesianIndex" in tn:
            n = f["size"][0]
            cells = np.frombuffer(f["data"], dtype="<i8").reshape(n, 2)
            if n == 50:
                food = cells
            else:
                snake = cells
        elif f.get("tag") == "array" and tn == "Core.Int64" and list(f["size"]) == [10, 10]:
            board = np.frombuffer(f["data"], dtype="<i8").reshape(10, 10).T  # column-major -> [r][c]
        elif f.get("tag") == "struct" and tn and "Xoshiro" in tn:
            rng = [deref(w) for w in f["data"]]
    assert food is not None and rng is not None
    scal = [(i, f) for i, f in enumerate(fields) if isinstance(f, (int, bool))]
    out = {
        "source": "trainers/very_long_training1.bson :: tr.game",
        "food_list_rc_1based": food.tolist(),
        # 5 words: s0..s3 and Julia 1.10's s4 (= s0+3s1+5s2+7s3 at seeding time)
        "xoshiro_state_after_draws_hex": ["0x%016x" % int.from_bytes(w["data"], "little") for w in rng],
    }
    with open(os.path.join(OUT, "g1_food_list.json"), "w") as f:
        json.dump(out, f, indent=1)
    g4 = {
        "source": "trainers/very_long_training1.bson :: tr.game (older 14-field struct)",
        "board_rc": board.tolist() if board is not None else None,
        "snake_rc_1based": snake.tolist() if snake is not None else None,
        "scalar_fields": [[i, (int(v) if not isinstance(v, bool) else v)] for i, v in scal],
    }
    with open(os.path.join(OUT, "g4_bson_game.json"), "w") as f:
        json.dump(g4, f)
    print("G1 food list:", food.tolist()[:5], "...", "state", out["xoshiro_state_after_draws_hex"])


# ----------------------------------------------------------------------------- GIF
def decode_gif(name):
    im = Image.open(os.path.join(REF, "trainer_gifs", name))
    x0, x1, y0, y1 = 131, 490, 12, 370  # board area of the 600x400 Plots.jl frame
    cw, ch = (x1 - x0 + 1) / 10.0, (y1 - y0 + 1) / 10.0
    boards = np.zeros((im.n_frames, 10, 10), dtype=np.int8)
    for k in range(im.n_frames):
        im.seek(k)
        fr = np.array(im.convert("RGB")).astype(int)
        for r in range(10):
            for c in range(10):
                # majority vote over a 9-pixel patch around the cell centre
                cy, cx = int(y0 + ch * (r + 0.5)), int(x0 + cw * (c + 0.5))
                px = fr[cy - 4:cy + 5:4, cx - 4:cx + 5:4].reshape(-1, 3).mean(0)
                R, G, B = px
                if R < 80 and G < 80 and B < 80:
                    v = -1          # black  = wall
                elif G > 150 and R < 120 and B < 120:
                    v = 1           # green  = snake
                elif R > 150 and G < 120 and B < 120:
                    v = 2           # red    = food
                elif R > 200 and G > 200 and B > 200:
                    v = 0           # white  = empty
                else:
                    raise ValueError("frame %d cell %d,%d colour %s" % (k, r, c, px))
                boards[k, r, c] = v
    return boards


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    golden_from_bson()
    b2 = decode_gif("very_long_double_training3.gif")
    np.save(os.path.join(OUT, "g2_boards_double3.npy"), b2)
    b5 = decode_gif("very_long_training1.gif")
    np.save(os.path.join(OUT, "g5_boards_training1.npy"), b5)
    print("G2 frames", b2.shape, "G5 frames", b5.shape)
    print(b2[0])
    print(b2[-1])


if __name__ == "__main__":
    main()
