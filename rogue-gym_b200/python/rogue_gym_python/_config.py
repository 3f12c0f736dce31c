"""GameConfig::to_json as used by GameState.dump_config (python/src/lib.rs:250-254,
core/src/lib.rs:147-149): serde's `skip_serializing_if = is_default` drops top-level sections
that equal their default, a section that differs is written in full, and `hide_dungeon` is
always written (core/src/lib.rs:44-86). Host-side only; nothing here touches game state."""
import json

_DUNGEON_DEFAULT = {
    "style": "rogue", "room_num_x": 3, "room_num_y": 3, "min_room_size": {"x": 4, "y": 4}, "enable_trap": True,
    "max_empty_rooms": 3, "amulet_level": 25, "maze_rate_inv": 15, "dark_level": 10,
    "hidden_passage_rate_inv": 40, "locked_door_rate_inv": 5, "max_extra_edges": 5, "door_unlock_rate_inv": 5,
    "passage_unlock_rate_inv": 3,
}
_EXPS = [10, 20, 40, 80, 160, 320, 640, 1300, 2600, 5200, 13000, 26000, 50000, 100000, 200000, 400000, 800000,
         2000000, 4000000, 8000000, 0xFFFFFFFF]
_INIT_ITEMS = [
    {"Noinit": {"kind": "Gold", "how_many": 0, "attr": 4}},
    {"Noinit": {"kind": {"Food": "Ration"}, "how_many": 1, "attr": 4}},
    {"Armor": {"name": "ring mail", "def_plus": 1}},
    {"Weapon": {"name": "mace", "num_plus": 0, "hit_plus": 1, "dam_plus": 1}},
    {"Weapon": {"name": "bow", "num_plus": 0, "hit_plus": 1, "dam_plus": 0}},
    {"Weapon": {"name": "arrow", "num_plus": 25, "hit_plus": 0, "dam_plus": 0}},
]
_PLAYER_DEFAULT = {"exps": _EXPS, "hunger_time": 1300, "init_hp": 12, "init_str": 16, "max_items": 27,
                   "init_items": _INIT_ITEMS, "heal_threshold": 20}
_GOLD_DEFAULT = {"rate_inv": 2, "base": 50, "per_level": 10, "minimum": 2}
_ENEMIES_DEFAULT = list(range(26))


def _filled(user, default):
    out = dict(default)
    out.update(user or {})
    return out


def dump_config(config_str, seed_override=None):
    cfg = json.loads(config_str) if config_str else {}
    out = {}
    if cfg.get("width", 80) != 80:
        out["width"] = cfg["width"]
    if cfg.get("height", 24) != 24:
        out["height"] = cfg["height"]
    seed = seed_override if seed_override is not None else cfg.get("seed")
    if seed is not None:
        out["seed"] = seed
    if cfg.get("seed_range") is not None:
        out["seed_range"] = cfg["seed_range"]
    dungeon = _filled(cfg.get("dungeon"), _DUNGEON_DEFAULT)
    if dungeon != _DUNGEON_DEFAULT:
        out["dungeon"] = dungeon
    item = cfg.get("item") or {}
    armor = {"armors": item.get("armor", {}).get("armors", list(range(8)))}
    for k, d in (("cursed_rate", 20), ("powerup_rate", 8)):
        if item.get("armor", {}).get(k, d) != d:
            armor[k] = item["armor"][k]
    weapon = {"weapons": item.get("weapon", {}).get("weapons", list(range(9)))}
    for k, d in (("cursed_rate", 10), ("powerup_rate", 5)):
        if item.get("weapon", {}).get(k, d) != d:
            weapon[k] = item["weapon"][k]
    gold = _filled(item.get("gold"), _GOLD_DEFAULT)
    if (armor, gold, weapon) != ({"armors": list(range(8))}, _GOLD_DEFAULT, {"weapons": list(range(9))}):
        out["item"] = {"armor": armor, "gold": gold, "weapon": weapon}
    if "keymap" in cfg:
        out["keymap"] = cfg["keymap"]
    player = _filled(cfg.get("player"), _PLAYER_DEFAULT)
    if player != _PLAYER_DEFAULT:
        out["player"] = player
    en = cfg.get("enemies") or {}
    enemies = {"enemies": en.get("enemies", _ENEMIES_DEFAULT)}
    for k, d in (("appear_rate_gold", 80), ("appear_rate_nogold", 25)):
        if en.get(k, d) != d:
            enemies[k] = en[k]
    if enemies != {"enemies": _ENEMIES_DEFAULT}:
        out["enemies"] = enemies
    out["hide_dungeon"] = cfg.get("hide_dungeon", True)
    return json.dumps(out, indent=2)
