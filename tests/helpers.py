"""Shared helpers of the GPU parity tests: canonical state dumps of both sides and diffs."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (shared constants are here, not in conftest.py: tests/golden/ref_tests has a conftest of its own, and `conftest` as a
# module name resolves to whichever was imported last)
MINI = {"width": 32, "height": 16, "seed": 4,
        "dungeon": {"style": "rogue", "room_num_x": 2, "room_num_y": 2, "min_room_size": {"x": 4, "y": 4}}}
MINI_NOMON = dict(MINI, enemies={"enemies": []})
DEFAULT = {}
CONFIGS = {"default": DEFAULT, "mini": MINI, "mini_nomon": MINI_NOMON,
           "default_clear": {"hide_dungeon": False, "enemies": {"enemies": []}},
           "wide": {"width": 160, "height": 48},
           "odd": {"width": 50, "height": 19, "dungeon": {"style": "rogue", "room_num_x": 2, "room_num_y": 2}},
           "deep": {"dungeon": {"style": "rogue", "dark_level": 2, "maze_rate_inv": 2, "hidden_passage_rate_inv": 4,
                                "locked_door_rate_inv": 2, "amulet_level": 1}},
           "grid4": {"width": 128, "height": 40, "dungeon": {"style": "rogue", "room_num_x": 4, "room_num_y": 4,
                                                           "max_empty_rooms": 6}}}


SCALAR_FIELDS = ("level", "px", "py", "hp", "hp_max", "exp", "plevel", "food_left", "quiet", "gold", "ui_dead", "steps",
                 "is_terminal", "message", "error", "n_monsters", "n_items", "n_cache", "status", "rng")
KEYS19 = np.frombuffer(b".hjklnbuy>sHJKLNBUY", np.uint8)


def gpu_dump(batch, env, with_maps=True):
    """rg_dump_env of one env of a rogue_gym_python._rogue_gym._Batch."""
    from rogue_gym_python import _cabi
    Cc = batch.C
    surface = np.zeros(Cc, np.uint8)
    attr = np.zeros(Cc, np.uint8)
    mons = np.zeros((_cabi.MAX_ROOMS, 8), np.int32)
    items = np.zeros((_cabi.MAX_ROOMS, 3), np.int32)
    cxy = np.zeros((_cabi.DIST_CACHE, 2), np.int32)
    maps = np.zeros((_cabi.DIST_CACHE, Cc), np.uint16) if with_maps else None
    rooms = np.zeros((_cabi.MAX_ROOMS, 8), np.int32)
    d = _cabi.Dump()
    d.surface, d.attr, d.monsters, d.items = surface.ctypes.data, attr.ctypes.data, mons.ctypes.data, items.ctypes.data
    d.cache_xy, d.rooms = cxy.ctypes.data, rooms.ctypes.data
    d.cache_maps = maps.ctypes.data if with_maps else None
    _cabi.check(batch.L.rg_dump_env(batch.h, env, C.byref(d)), batch.h)
    s = d.s.as_dict()
    nrooms = batch.params.room_num_x * batch.params.room_num_y
    return dict(scalars=s, surface=surface, attr=attr & 0x7F, monsters=mons[: s["n_monsters"]], items=items[: s["n_items"]],
                cache_xy=cxy[: s["n_cache"]], cache_maps=(maps[: s["n_cache"]] if with_maps else None), rooms=rooms[:nrooms])


def oracle_dump(env, with_maps=True):
    s = env.scalars().as_dict()
    surface, attr = env.grid()
    mons, items = env.entities()
    cxy, maps = env.dist_cache(with_maps)
    return dict(scalars=s, surface=surface.ravel(), attr=attr.ravel(), monsters=mons, items=items, cache_xy=cxy,
                cache_maps=maps, rooms=env.rooms())


def diff_dumps(g, o, W):
    """Returns a list of human-readable differences (empty = identical)."""
    out = []
    if o["scalars"]["error"] in (3, 4) or g["scalars"]["error"] in (3, 4):
        # a panicked env stops mid-step on both sides; only the error itself is defined
        if o["scalars"]["error"] != g["scalars"]["error"]:
            out.append("error: gpu %d oracle %d" % (g["scalars"]["error"], o["scalars"]["error"]))
        return out
    for k in SCALAR_FIELDS:
        if g["scalars"][k] != o["scalars"][k]:
            out.append("scalar %s: gpu %s oracle %s" % (k, g["scalars"][k], o["scalars"][k]))
    for k in ("surface", "attr"):
        bad = np.nonzero(g[k] != o[k])[0]
        if len(bad):
            out.append("%s differs at %d cells, first (x,y,gpu,oracle): %s" % (
                k, len(bad), [(int(i % W), int(i // W), int(g[k][i]), int(o[k][i])) for i in bad[:6]]))
    for k in ("monsters", "items", "cache_xy", "rooms"):
        if g[k].shape != o[k].shape or not np.array_equal(g[k], o[k]):
            out.append("%s: gpu %s oracle %s" % (k, g[k].tolist(), o[k].tolist()))
    if g["cache_maps"] is not None and o["cache_maps"] is not None and g["cache_maps"].shape == o["cache_maps"].shape:
        for i in range(len(g["cache_maps"])):
            bad = np.nonzero(g["cache_maps"][i] != o["cache_maps"][i])[0]
            if len(bad):
                out.append("dist map %d differs at %d cells, first (x,y,gpu,oracle): %s" % (
                    i, len(bad), [(int(j % W), int(j // W), int(g["cache_maps"][i][j]), int(o["cache_maps"][i][j]))
                                  for j in bad[:6]]))
    return out


def render(screen_row, W):
    raw = bytes(screen_row)
    return "\n".join(raw[i:i + W].decode("latin-1") for i in range(0, len(raw), W))


def diff_obs(batch, ob, step, W, live=None):
    """Compares the host mirrors of a _Batch with OracleBatch.obs(); returns a list of differences."""
    out = []
    n = batch.n
    if live is None:
        live = np.ones(n, bool)
    pairs = (("screen", batch.screen, ob["screen"]), ("history", batch.history, ob["history"]),
             ("status", batch.status, ob["status"]), ("message", batch.message, ob["message"]),
             ("done", batch.done, ob["done"]))
    for name, g, o in pairs:
        neq = (g != o)
        if neq.ndim > 1:
            neq = neq.any(axis=1)
        neq &= live
        if neq.any():
            i = int(np.nonzero(neq)[0][0])
            msg = "step %d: %s differs for %d envs, first env %d" % (step, name, int(neq.sum()), i)
            if name in ("screen",):
                msg += "\n--- gpu\n%s\n--- oracle\n%s" % (render(g[i], W), render(o[i], W))
            elif name == "history":
                bad = np.nonzero(g[i] != o[i])[0]
                msg += " cells %s" % [(int(j % W), int(j // W), int(g[i][j]), int(o[i][j])) for j in bad[:8]]
            else:
                msg += ": gpu %s oracle %s" % (g[i].tolist(), o[i].tolist())
            out.append(msg)
    return out


# ---- the reference's rendered recordings (tests/golden/reference_gif_frames.json, made by make_gif_golden.py)
def status_line(st, width):
    """impl Display for Status (core/src/character/player.rs:433-449), cut at the screen width like the recording."""
    hunger = {0: "", 1: "Hungry", 2: "Weak"}.get(int(st[9]), "")
    s = "Level: %2d Gold: %5d Hp: %2d(%2d) Str: %2d(%2d) Arm: %2d Exp: %2d/%2d " % tuple(int(v) for v in st[:9]) + hunger
    return s[:width].ljust(width)


def check_against_recording(react, screen, status, msg_flags, keys, frames):
    """act2gif (act2gif/src/draw.rs:44-68) emits one frame per Reaction::Redraw; the terminal it renders is persistent.
    `react(key)` steps the game under test, `screen()` gives its rows, `status()` its 10-vector, `msg_flags()` the
    PlayerState message bits of the last action. Checks, frame by frame:
      * rows 1 .. H-2 (the dungeon) equal the game's screen after the action that produced the frame; an action that
        produced no frame (a blocked move) must have left the screen as it was;
      * the status line is the status before or after that action (StatusUpdated comes after Redraw in the reaction list,
        so the line lags by one frame), blank before the first update;
      * the messages shown, in order, are the ones the game's events call for (gold pick-ups, secret door, no downstair).
    Returns the number of frames matched."""
    H, W = len(frames[0]), len(frames[0][0])
    fi, shown, events = 0, [], []
    prev_screen, prev_status = screen(), [int(v) for v in status()]
    blank = " " * W
    for k in keys:
        react(k)
        scr, st, mf = screen(), [int(v) for v in status()], int(msg_flags())
        if st[1] > prev_status[1]:
            events.append(("You got %d Gold" % (st[1] - prev_status[1]))[:W])
        if mf & 32:  # MessageFlag SECRET_DOOR (python/src/flags.rs)
            events.append("You found a secret door"[:W])
        if mf & 64:  # NO_DOWNSTAIR
            events.append("Hmm... there seems to be no downstair"[:W])
        if fi < len(frames) and frames[fi][1:H - 1] == scr[1:H - 1] and (scr != prev_screen or k in b"s>" or fi == 0):
            fr = frames[fi]
            assert fr[H - 1] in (blank, status_line(prev_status, W), status_line(st, W)), (fi, fr[H - 1])
            if fr[0] != blank and (not shown or shown[-1] != fr[0].rstrip()):
                shown.append(fr[0].rstrip())
            fi += 1
        else:
            assert scr == prev_screen, "action %r changed the screen but matches no frame (frame %d)" % (chr(k), fi)
        prev_screen, prev_status = scr, st
    # every message that was displayed was called for, in order (the last event may have had no later frame to show it)
    it = iter(events)
    assert all(any(m == e for e in it) for m in shown), (shown, events)
    return fi
