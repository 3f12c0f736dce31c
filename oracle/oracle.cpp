// oracle/oracle.cpp — CPU oracle for the rogue-gym hot path.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h). A deliberately plain, serial restatement of
// the reference's algorithm that keeps the reference's own data-structure shapes
// (ordered sets with nth-select, ordered maps keyed by [level,x,y], reaction lists) so
// that each function can be read next to the Rust it follows. Citations are relative
// to /root/reference (kngwyu/rogue-gym @ c78608b).
//
// Third-party arithmetic restated here (crates are not vendored in the reference):
//   rand_xorshift 0.2  XorShiftRng        -> Rng::next_u32 / next_u64 / seed
//   rand 0.7           UniformInt::sample_single, SliceRandom::choose -> Rng::range32/64
//   rect-iter 0.3      RectRange<i32>     -> Rect
//   enum-iterator 0.6  Direction order    -> DX/DY tables
#include "oracle.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Panic {
  std::string msg;
};
struct GameError {
  int code;
  std::string msg;
};

[[noreturn]] void panic(const char* m) { throw Panic{m}; }

// ---------------------------------------------------------------- RNG (core/src/rng.rs)
struct Rng {
  uint32_t x = 0, y = 0, z = 0, w = 0;
  uint64_t draws = 0;
  // rng.rs:47-55: the u128 seed is transmuted to 16 LE bytes, XorShiftRng::from_seed reads
  // four LE u32 words; an all-zero seed is replaced (rand_xorshift 0.2, unpinned).
  void seed(uint64_t lo, uint64_t hi) {
    x = (uint32_t)lo;
    y = (uint32_t)(lo >> 32);
    z = (uint32_t)hi;
    w = (uint32_t)(hi >> 32);
    if ((x | y | z | w) == 0) x = y = z = w = 0x0BAD5EEDu;
    draws = 0;
  }
  uint32_t next_u32() {
    ++draws;
    uint32_t t = x ^ (x << 11);
    x = y;
    y = z;
    z = w;
    w = w ^ (w >> 19) ^ (t ^ (t >> 8));
    return w;
  }
  uint64_t next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return (hi << 32) | lo;
  }
  // rand 0.7 UniformInt::sample_single, 32-bit lane (u32 and i32 call sites)
  uint32_t range32(uint32_t lo, uint32_t hi, bool is_signed = false) {
    bool ok = is_signed ? ((int32_t)lo < (int32_t)hi) : (lo < hi);
    if (!ok) panic("invalid range!!");  // rng.rs:87
    uint32_t range = hi - lo;
    uint32_t zone = (range << __builtin_clz(range)) - 1u;
    for (;;) {
      uint32_t v = next_u32();
      uint64_t m = (uint64_t)v * (uint64_t)range;
      if ((uint32_t)m <= zone) return lo + (uint32_t)(m >> 32);
    }
  }
  int32_t range_i32(int32_t lo, int32_t hi) { return (int32_t)range32((uint32_t)lo, (uint32_t)hi, true); }
  // 64-bit lane (usize, u64, i64 call sites)
  uint64_t range64(uint64_t lo, uint64_t hi, bool is_signed = false) {
    bool ok = is_signed ? ((int64_t)lo < (int64_t)hi) : (lo < hi);
    if (!ok) panic("invalid range!!");
    uint64_t range = hi - lo;
    uint64_t zone = (range << __builtin_clzll(range)) - 1ull;
    for (;;) {
      uint64_t v = next_u64();
      unsigned __int128 m = (unsigned __int128)v * range;
      if ((uint64_t)m <= zone) return lo + (uint64_t)(m >> 64);
    }
  }
  int64_t range_i64(int64_t lo, int64_t hi) { return (int64_t)range64((uint64_t)lo, (uint64_t)hi, true); }
  size_t range_usize(size_t lo, size_t hi) { return (size_t)range64(lo, hi); }
  // rng.rs:91-93
  bool does_happen(uint32_t p_inv) { return range32(0, p_inv) == 0; }
  // rng.rs:95-98
  bool parcent(uint32_t p) { return range32(1, 101) <= p; }
};

// ------------------------------------------------------------- FenwickSet (fenwick.rs)
// Ordered integer set over [0, cap) with nth-smallest and uniform select. The Fenwick
// tree of the reference is an implementation detail; only set semantics are observable.
struct OrdSet {
  std::vector<uint8_t> bits;
  size_t count = 0;
  OrdSet() = default;
  explicit OrdSet(size_t cap) : bits(cap, 0) {}
  static OrdSet from_range(size_t lo, size_t hi) {
    OrdSet s(hi);
    for (size_t i = lo; i < hi; ++i) s.insert(i);
    return s;
  }
  bool contains(size_t e) const { return e < bits.size() && bits[e]; }
  bool insert(size_t e) {  // fenwick.rs:48-56
    if (e >= bits.size() || bits[e]) return false;
    bits[e] = 1;
    ++count;
    return true;
  }
  bool remove(size_t e) {  // fenwick.rs:60-68
    if (e >= bits.size() || !bits[e] || count == 0) return false;
    bits[e] = 0;
    --count;
    return true;
  }
  std::optional<size_t> nth(size_t n) const {  // fenwick.rs:77-84
    size_t seen = 0;
    for (size_t i = 0; i < bits.size(); ++i)
      if (bits[i]) {
        if (seen == n) return i;
        ++seen;
      }
    return std::nullopt;
  }
  size_t len() const { return count; }
  std::optional<size_t> select(Rng& rng) const {  // fenwick.rs:90-96 (usize -> 64-bit lane)
    if (count == 0) return std::nullopt;
    size_t n = rng.range_usize(0, count);
    return nth(n);
  }
};

// --------------------------------------------------------------- geometry (coord.rs)
struct Coord {
  int x = 0, y = 0;
  bool operator==(const Coord& o) const { return x == o.x && y == o.y; }
  bool operator!=(const Coord& o) const { return !(*this == o); }
  bool operator<(const Coord& o) const { return x != o.x ? x < o.x : y < o.y; }
};
enum Dir { Up = 0, Down, Left, Right, LeftUp, RightUp, LeftDown, RightDown, Stay };
const int DX[9] = {0, 0, -1, 1, -1, 1, -1, 1, 0};   // coord.rs:227-241
const int DY[9] = {-1, 1, 0, 0, -1, -1, 1, 1, 0};
inline Coord step(Coord c, int d) { return Coord{c.x + DX[d], c.y + DY[d]}; }
inline bool is_diag(int d) { return d >= LeftUp && d <= RightDown; }
inline int reverse(int d) {  // coord.rs:272-285
  static const int R[9] = {Down, Up, Right, Left, RightDown, LeftDown, RightUp, LeftUp, Stay};
  return R[d];
}

// rect-iter RectRange<i32>: half-open, row-major
struct Rect {
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  static Rect from_corners(Coord a, Coord b) {
    if (a.x >= b.x || a.y >= b.y) panic("RectRange::from_corners on empty range");
    return Rect{a.x, a.y, b.x, b.y};
  }
  bool is_valid() const { return x0 < x1 && y0 < y1; }
  int xlen() const { return x1 - x0; }
  int ylen() const { return y1 - y0; }
  size_t len() const { return (size_t)xlen() * (size_t)ylen(); }
  bool contains(Coord c) const { return c.x >= x0 && c.x < x1 && c.y >= y0 && c.y < y1; }
  std::optional<size_t> index(Coord c) const {
    if (!contains(c)) return std::nullopt;
    return (size_t)(c.x - x0) + (size_t)(c.y - y0) * (size_t)xlen();
  }
  std::optional<Coord> nth(size_t n) const {
    if (n >= len()) return std::nullopt;
    return Coord{x0 + (int)(n % xlen()), y0 + (int)(n / xlen())};
  }
  bool is_vert_edge(Coord c) const { return c.x == x0 || c.x == x1 - 1; }
  bool is_horiz_edge(Coord c) const { return c.y == y0 || c.y == y1 - 1; }
  bool is_edge(Coord c) const { return is_vert_edge(c) || is_horiz_edge(c); }
};

// ------------------------------------------------------------- field (field.rs, tile.rs)
enum Surface : uint8_t { SPassage = 0, SFloor, SWallX, SWallY, SStair, SDoor, STrap, SNone };  // rogue/mod.rs:137-146
inline uint8_t surface_tile(uint8_t s) {  // rogue/mod.rs:148-161
  static const char T[8] = {'#', '.', '-', '|', '%', '+', '^', ' '};
  return (uint8_t)T[s];
}
inline bool can_walk(uint8_t s) { return !(s == SWallX || s == SWallY || s == SNone); }  // :176-183
enum : uint32_t { IS_VISITED = 1, IS_HIDDEN = 2, IS_VISIBLE = 4, HAS_DRAWN = 8, IS_LOCKED = 16, IS_DARK = 32 };

struct Cell {
  uint8_t surface = SNone;
  uint32_t attr = 0;
  void visible(bool on) { on ? (attr |= IS_VISIBLE) : (attr &= ~IS_VISIBLE); }
  void approached() {  // field.rs:20-26
    if (attr & IS_HIDDEN) return;
    attr |= HAS_DRAWN;
    visible(true);
  }
  void left() {  // field.rs:30-34
    if (attr & IS_DARK) visible(false);
  }
  bool is_obj_visible() const { return (attr & IS_VISIBLE) || (attr & HAS_DRAWN); }
  bool is_hidden() const { return attr & IS_HIDDEN; }
  bool is_locked() const { return attr & IS_LOCKED; }
  void unlock() {  // field.rs:84-87
    attr &= ~(IS_LOCKED | IS_HIDDEN);
    visible(true);
  }
  uint8_t tile() const { return (attr & IS_VISIBLE) ? surface_tile(surface) : (uint8_t)' '; }  // field.rs:92-98
};

struct Field {
  int w = 0, h = 0;
  std::vector<Cell> cells;
  Field() = default;
  Field(int w_, int h_) : w(w_), h(h_), cells((size_t)w_ * h_) {}
  // field.rs:161-175: the bounds test is `>` (accepts x == w, y == h); reproduced, and an
  // index past the end is a Rust panic.
  Cell* try_get(Coord c) {
    if (c.x < 0 || c.y < 0) return nullptr;
    if (c.x > w || c.y > h) return nullptr;
    size_t id = (size_t)c.y * w + c.x;
    if (id >= cells.size()) panic("Field index out of bounds");
    return &cells[id];
  }
  const Cell* try_get(Coord c) const { return const_cast<Field*>(this)->try_get(c); }
};

// ---------------------------------------------------------------- rooms (rooms.rs)
enum RoomKind { KNormal = 0, KMaze = 1, KEmpty = 2 };
struct Room {
  int kind = KNormal;
  Rect range;            // Normal: room rect; Maze: maze.range
  OrdSet maze_passages;  // Maze only (maze.rs:13-14)
  Coord up_left;         // Empty only
  bool is_dark = false;
  size_t id = 0;
  Rect assigned;
  bool is_visited = false;
  bool has_gold = false;
  OrdSet empty_cells, nocharacter_cells;

  bool has_range() const { return kind != KEmpty; }  // rooms.rs:84-90
  bool contains(Coord c) const { return assigned.contains(c); }
  std::optional<size_t> cell_id(Coord c) const {
    if (!has_range()) return std::nullopt;
    return range.index(c);
  }
  bool fill_cell(Coord c, bool is_character) {  // rooms.rs:96-105
    auto id_ = cell_id(c);
    if (!id_) return false;
    if (is_character) nocharacter_cells.remove(*id_);
    return empty_cells.remove(*id_);
  }
  bool unfill_cell(Coord c, bool is_character) {  // rooms.rs:107-116
    auto id_ = cell_id(c);
    if (!id_) return false;
    if (is_character) nocharacter_cells.insert(*id_);
    return empty_cells.insert(*id_);
  }
  std::optional<Coord> select_cell(Rng& rng, bool is_character) const {  // rooms.rs:132-144
    if (!has_range()) return std::nullopt;
    const OrdSet& set = is_character ? nocharacter_cells : empty_cells;
    auto n = set.select(rng);
    if (!n) return std::nullopt;
    return range.nth(*n);
  }
  bool maze_has_cd(Coord c) const {  // maze.rs:25-31
    auto i = range.index(c);
    return i && maze_passages.contains(*i);
  }
};

// maze.rs:38-89
void dig_maze_impl(const Rect& range, Rng& rng, const std::function<void(Coord)>& reg, std::set<Coord>& used,
                   Coord cur) {
  for (;;) {
    int pick = -1;
    uint32_t k = 0;
    for (int d = 0; d < 4; ++d) {
      Coord nxt{cur.x + DX[d] * 2, cur.y + DY[d] * 2};
      if (!(range.contains(nxt) && !used.count(nxt))) continue;
      if (rng.does_happen(k + 1)) pick = d;  // reservoir: every candidate draws
      ++k;
    }
    if (pick < 0) break;
    Coord c = cur;
    for (int i = 0; i < 2; ++i) {
      c = step(c, pick);
      if (used.insert(c).second) reg(c);
    }
    dig_maze_impl(range, rng, reg, used, Coord{cur.x + DX[pick] * 2, cur.y + DY[pick] * 2});
  }
}
void dig_maze(const Rect& range, Rng& rng, const std::function<void(Coord)>& reg) {
  Coord start{range.x0, range.y0};
  reg(start);
  std::set<Coord> used;
  used.insert(start);
  dig_maze_impl(range, rng, reg, used, start);
}

OrdSet gen_empty_cells(const Room& r) {  // rooms.rs:147-162
  if (r.kind == KNormal) {
    OrdSet s(r.range.len());
    for (size_t i = 0; i < r.range.len(); ++i) {
      Coord c = *r.range.nth(i);
      if (!r.range.is_edge(c)) s.insert(i);
    }
    return s;
  }
  if (r.kind == KMaze) return r.maze_passages;
  return OrdSet(1);
}

Room make_room(bool is_empty, Coord room_size, Coord lower_left, size_t id, const orc_params& cfg, uint32_t level,
               Rng& rng) {  // rooms.rs:214-269
  Room r;
  r.id = id;
  r.assigned = Rect::from_corners(lower_left, Coord{lower_left.x + room_size.x, lower_left.y + room_size.y});
  if (is_empty) {
    int x = rng.range_i32(1, room_size.x - 1) + lower_left.x;
    int y = rng.range_i32(1, room_size.y - 1) + lower_left.y;
    r.kind = KEmpty;
    r.up_left = Coord{x, y};
    r.is_dark = true;
  } else {
    r.is_dark = rng.range32(0, cfg.dark_level) < level;
    if (r.is_dark && rng.does_happen(cfg.maze_rate_inv)) {
      r.kind = KMaze;
      r.range = Rect::from_corners(lower_left, Coord{lower_left.x + room_size.x - 1, lower_left.y + room_size.y - 1});
      r.maze_passages = OrdSet(r.range.len());
      Rect range = r.range;
      OrdSet* ps = &r.maze_passages;
      dig_maze(range, rng, [&](Coord c) {
        auto i = range.index(c);
        if (!i) panic("dig_maze produced invalid Coordinate!");
        ps->insert(*i);
      });
    } else {
      r.kind = KNormal;
      int sx = rng.range_i32(cfg.min_room_x, room_size.x);
      int sy = rng.range_i32(cfg.min_room_y, room_size.y);
      int lx = rng.range_i32(0, room_size.x - sx) + lower_left.x;
      int ly = rng.range_i32(0, room_size.y - sy) + lower_left.y;
      r.range = Rect::from_corners(Coord{lx, ly}, Coord{lx + sx, ly + sy});
    }
  }
  r.empty_cells = gen_empty_cells(r);
  r.nocharacter_cells = r.empty_cells;
  return r;
}

std::vector<Room> gen_rooms(uint32_t level, const orc_params& cfg, Rng& rng) {  // rooms.rs:165-211
  int nx = cfg.room_num_x, ny = cfg.room_num_y;
  size_t room_num = (size_t)(nx * ny);
  Coord room_size{cfg.width / nx, cfg.height / ny};
  uint32_t empty_num = rng.range32(0, cfg.max_empty_rooms + 1);
  if (empty_num >= (uint32_t)room_num) empty_num = (uint32_t)room_num - 1;
  std::vector<uint8_t> empty_rooms(room_num, 0);
  {  // RngHandle::select(0..room_num).take(empty_num)  rng.rs:62-76,129-143
    OrdSet rest = OrdSet::from_range(0, room_num);
    for (uint32_t i = 0; i < empty_num; ++i) {
      size_t num_rests = rest.len();
      if (num_rests == 0) break;
      size_t n = rng.range_usize(0, num_rests);
      size_t res = *rest.nth(n);
      rest.remove(res);
      empty_rooms[res] = 1;
    }
  }
  std::vector<Room> rooms;
  size_t i = 0;
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x, ++i) {
      Coord rs = room_size;
      Coord ll;
      if (y == 0) {
        rs.y -= 1;
        ll = Coord{rs.x * x, rs.y * y + 1};
      } else {
        ll = Coord{rs.x * x, rs.y * y};
      }
      if (ll.y + rs.y == cfg.height) rs.y -= 1;
      rooms.push_back(make_room(empty_rooms[i], rs, ll, i, cfg, level, rng));
    }
  return rooms;
}

// ------------------------------------------------------------ passages (passages.rs)
struct Positioned {
  Coord cd;
  uint8_t surface;
};

std::vector<Coord> edges(const Rect& r, int direction, bool inclusive) {  // passages.rs:181-219
  int off = inclusive ? 1 : 0;
  int bound_x = r.x1 - off, bound_y = r.y1 - off;
  std::vector<Coord> out;
  switch (direction) {
    case Down:
      for (Coord c{r.x0 + off, r.y1 - 1}; c.x < bound_x; ++c.x) out.push_back(c);
      break;
    case Left:
      for (Coord c{r.x0, r.y0 + off}; c.y < bound_y; ++c.y) out.push_back(c);
      break;
    case Right:
      for (Coord c{r.x1 - 1, r.y0 + off}; c.y < bound_y; ++c.y) out.push_back(c);
      break;
    case Up:
      for (Coord c{r.x0 + off, r.y0}; c.x < bound_x; ++c.x) out.push_back(c);
      break;
    default:
      panic("[passages::connet_2rooms] invalid direction");
  }
  return out;
}

Coord choose(const std::vector<Coord>& v, Rng& rng) {  // rand 0.7 SliceRandom::choose (usize lane)
  return v[rng.range_usize(0, v.size())];
}

Coord select_start_or_end(const Room& room, int direction, Rng& rng) {  // passages.rs:143-179
  if (room.kind == KNormal) {
    auto e = edges(room.range, direction, true);
    if (e.empty()) panic("choose on empty edge");
    return choose(e, rng);
  }
  if (room.kind == KMaze) {
    Rect range = room.range;
    while (range.is_valid()) {
      std::vector<Coord> cand;
      for (Coord c : edges(range, direction, false))
        if (room.maze_has_cd(c)) cand.push_back(c);
      if (!cand.empty()) return choose(cand, rng);
      switch (direction) {
        case Down: range.y1 -= 1; break;
        case Left: range.x0 -= 1; break;
        case Right: range.x1 -= 1; break;
        case Up: range.y0 -= 1; break;
        default: panic("unreachable");
      }
    }
    panic("cannot find maze floor in passages::select_start_or_end");
  }
  return room.up_left;
}

uint8_t door_kind(const Room& r) { return r.kind == KNormal ? SDoor : SPassage; }  // passages.rs:135-141

void connect_2rooms(const Room* room1, const Room* room2, int direction, Rng& rng,
                    std::vector<Positioned>& out) {  // passages.rs:84-133
  if (direction == Up || direction == Left) {
    std::swap(room1, room2);
    direction = reverse(direction);
  }
  Coord start = select_start_or_end(*room1, direction, rng);
  Coord end = select_start_or_end(*room2, reverse(direction), rng);
  out.push_back({start, door_kind(*room1)});
  out.push_back({end, door_kind(*room2)});
  Coord turn_start, turn_end;
  int turn_dir;
  if (direction == Down) {
    int y = rng.range_i32(start.y + 1, end.y);
    turn_dir = (start.x < end.x) ? Right : Left;
    turn_start = Coord{start.x, y};
    turn_end = Coord{end.x, y};
  } else if (direction == Right) {
    int x = rng.range_i32(start.x + 1, end.x);
    turn_dir = (start.y < end.y) ? Down : Up;
    turn_start = Coord{x, start.y};
    turn_end = Coord{x, end.y};
  } else {
    panic("unreachable");
  }
  bool first = true;
  for (Coord c = start; c != turn_start; c = step(c, direction)) {
    if (first) {
      first = false;  // .skip(1)
      continue;
    }
    out.push_back({c, SPassage});
  }
  for (Coord c = turn_start; c != turn_end; c = step(c, turn_dir)) out.push_back({c, SPassage});
  for (Coord c = turn_end; c != end; c = step(c, direction)) out.push_back({c, SPassage});
}

struct Node {  // passages.rs:244-270
  std::vector<uint8_t> connections;
  std::map<size_t, int> candidates;
};

std::optional<std::pair<size_t, int>> select_candidate(size_t num_rooms, const Node& node, Rng& rng,
                                                       const std::function<bool(size_t)>& pred) {  // :69-82
  std::optional<std::pair<size_t, int>> last;
  uint32_t k = 0;
  for (size_t i = 0; i < num_rooms; ++i) {
    if (!pred(i)) continue;
    auto it = node.candidates.find(i);
    if (it == node.candidates.end()) continue;
    if (rng.does_happen(k + 1)) last = std::make_pair(i, it->second);
    ++k;
  }
  return last;
}

void dig_passages(const std::vector<Room>& rooms, int xrooms, int yrooms, Rng& rng, uint32_t max_extra_edges,
                  std::vector<Positioned>& out) {  // passages.rs:16-67
  size_t num_rooms = rooms.size();
  std::vector<Node> graph(num_rooms);
  {
    size_t i = 0;
    for (int y = 0; y < yrooms; ++y)
      for (int x = 0; x < xrooms; ++x, ++i) {
        graph[i].connections.assign(num_rooms, 0);
        for (int d = 0; d < 4; ++d) {
          int nx = x + DX[d], ny = y + DY[d];
          if (nx < 0 || ny < 0 || nx >= xrooms || ny >= yrooms) continue;
          graph[i].candidates[(size_t)(nx + ny * xrooms)] = d;
        }
      }
  }
  auto connect = [&](size_t a, size_t b) {
    graph[a].connections[b] = 1;
    graph[b].connections[a] = 1;
  };
  OrdSet selected(num_rooms);
  size_t cur_room = rng.range_usize(0, num_rooms);
  selected.insert(cur_room);
  while (selected.len() < num_rooms) {
    auto nxt = select_candidate(num_rooms, graph[cur_room], rng, [&](size_t id) { return !selected.contains(id); });
    if (nxt) {
      selected.insert(nxt->first);
      connect(cur_room, nxt->first);
      connect_2rooms(&rooms[cur_room], &rooms[nxt->first], nxt->second, rng, out);
    } else {
      cur_room = *selected.select(rng);
    }
  }
  uint32_t try_num = rng.range32(0, max_extra_edges);
  for (uint32_t t = 0; t < try_num; ++t) {
    size_t room1 = rng.range_usize(0, num_rooms);
    auto sel = select_candidate(num_rooms, graph[room1], rng,
                                [&](size_t id) { return !graph[room1].connections[id]; });
    if (sel) {
      connect(room1, sel->first);
      connect_2rooms(&rooms[room1], &rooms[sel->first], sel->second, rng, out);
    }
  }
}

// ------------------------------------------------------------------ enemies (enemies.rs)
enum : uint32_t { ATTR_MEAN = 1, ATTR_RANDOM = 0x200, ATTR_CONFUSED = 0x400 };
struct Enemy {
  int kind = 0;
  int64_t hp = 0, max_hp = 0, level = 0;
  int32_t defense = 0;
  uint32_t exp = 0;
  uint32_t attr = 0;
  bool running = false;
  bool is_mean() const { return attr & ATTR_MEAN; }
  bool is_random() const { return attr & ATTR_RANDOM; }
  bool is_confused() const { return attr & ATTR_CONFUSED; }
};
using Path = std::array<int, 3>;  // DungeonPath [level, x, y], lexicographic (dungeon/mod.rs:111-116)
using EnemyMap = std::map<Path, std::shared_ptr<Enemy>>;

enum MoveKind { CantMove = 0, CanMove = 1, Reach = 2 };
struct MoveResult {
  int kind;
  Path to;
};

// ------------------------------------------------------------------ floor (floor.rs)
struct Item {
  uint32_t amount;
};

// Damage for Dice<HitPoint>::random (character/mod.rs:229-235): `times` draws of 1..=max on the given stream
static int64_t dice_roll(Rng& rng, int times, int64_t mx) {
  int64_t acc = 0;
  for (int i = 0; i < times; ++i) acc += rng.range_i64(1, mx + 1);
  return acc;
}

struct Floor {
  std::vector<Room> rooms;
  std::set<Coord> doors;
  Field field;
  OrdSet non_empty_rooms;
  std::map<Coord, Item> items;  // HashMap in the reference; never iterated on the path

  static uint32_t gen_attr(uint8_t surface, bool is_dark, Rng& rng, uint32_t level, const orc_params& cfg) {  // :420-451
    uint32_t attr = 0;
    switch (surface) {
      case SPassage:
        if (rng.range32(0, cfg.dark_level) < level && rng.does_happen(cfg.hidden_passage_rate_inv)) attr |= IS_HIDDEN;
        break;
      case SDoor:
        if (rng.range32(0, cfg.dark_level) < level && rng.does_happen(cfg.locked_door_rate_inv)) attr |= IS_LOCKED;
        break;
      case SFloor:
        if (is_dark) attr |= IS_DARK;
        break;
      default:
        break;
    }
    return attr;
  }

  static Floor gen_floor(uint32_t level, const orc_params& cfg, Rng& rng) {  // floor.rs:50-104
    Floor f;
    f.rooms = gen_rooms(level, cfg, rng);
    f.field = Field(cfg.width, cfg.height);
    for (const Room& room : f.rooms) {  // Room::draw rooms.rs:58-82
      auto put = [&](Coord c, uint8_t s) {
        Cell* cell = f.field.try_get(c);
        if (!cell) throw GameError{ORC_ERR_SETTING, "Error in gen_floor"};
        cell->surface = s;
        cell->attr = gen_attr(s, room.is_dark, rng, level, cfg);
      };
      if (room.kind == KNormal) {
        for (size_t i = 0; i < room.range.len(); ++i) {
          Coord c = *room.range.nth(i);
          uint8_t s = room.range.is_horiz_edge(c) ? SWallX : room.range.is_vert_edge(c) ? SWallY : SFloor;
          put(c, s);
        }
      } else if (room.kind == KMaze) {
        for (size_t i = 0; i < room.maze_passages.bits.size(); ++i)
          if (room.maze_passages.bits[i]) put(*room.range.nth(i), SPassage);
      }
    }
    std::vector<Positioned> passages;
    dig_passages(f.rooms, cfg.room_num_x, cfg.room_num_y, rng, cfg.max_extra_edges, passages);
    for (const Positioned& p : passages) {
      if (p.surface == SDoor) f.doors.insert(p.cd);
      Cell* cell = f.field.try_get(p.cd);
      if (!cell) throw GameError{ORC_ERR_SETTING, "Floor::new dig_passges returned invalid index"};
      cell->attr = gen_attr(p.surface, false, rng, level, cfg);
      if (!cell->is_hidden() && !cell->is_locked()) cell->surface = p.surface;
    }
    f.non_empty_rooms = OrdSet(f.rooms.size());  // Floor::new floor.rs:29-46
    for (const Room& r : f.rooms)
      if (r.kind != KEmpty) f.non_empty_rooms.insert(r.id);
    return f;
  }

  std::optional<Coord> select_cell(Rng& rng, bool is_character) const {  // floor.rs:333-346
    OrdSet cand = non_empty_rooms;
    while (cand.len() > 0) {
      auto idx = cand.select(rng);
      if (!idx) panic("Logic Error in floor::select_cell");
      auto cd = rooms[*idx].select_cell(rng, is_character);
      if (cd) return cd;
      cand.remove(*idx);
    }
    return std::nullopt;
  }

  bool set_obj(Coord c, bool is_character) {  // floor.rs:315-321
    for (Room& r : rooms)
      if (r.contains(c)) return r.fill_cell(c, is_character);
    return false;
  }
  bool remove_obj(Coord c, bool is_character) {  // floor.rs:324-330
    for (Room& r : rooms)
      if (r.contains(c)) return r.unfill_cell(c, is_character);
    return false;
  }

  // floor.rs:169-182; nullopt <=> the `?` early-outs
  std::optional<bool> can_move_impl(Coord cd, int d, bool is_enemy) const {
    const Cell* nxt = field.try_get(step(cd, d));
    if (!nxt) return std::nullopt;
    bool res = can_walk(nxt->surface);
    if (!is_enemy) {
      res &= !nxt->is_hidden();
      res &= !nxt->is_locked();
    }
    if (is_diag(d)) {
      const Cell* cx = field.try_get(Coord{cd.x + DX[d], cd.y});
      if (!cx) return std::nullopt;
      res &= can_walk(cx->surface);
      const Cell* cy = field.try_get(Coord{cd.x, cd.y + DY[d]});
      if (!cy) return std::nullopt;
      res &= can_walk(cy->surface);
    }
    return res;
  }
  std::optional<Coord> can_move_player(Coord cd, int d) const {
    if (can_move_impl(cd, d, false).value_or(false)) return step(cd, d);
    return std::nullopt;
  }
  bool can_move_enemy(Coord cd, int d) const { return can_move_impl(cd, d, true).value_or(false); }

  std::optional<size_t> cd_to_room_id(Coord c) const {  // floor.rs:194-200
    for (size_t i = 0; i < rooms.size(); ++i)
      if (rooms[i].assigned.contains(c)) return i;
    return std::nullopt;
  }

  void with_current_room(Coord cd, const std::function<bool(Room&)>& select,
                         const std::function<void(Cell&, bool)>& mark) {  // floor.rs:201-228
    auto id = cd_to_room_id(cd);
    if (!id) throw GameError{ORC_ERR_SETTING, "[Floor::with_current_room] no room for given coord"};
    if (!select(rooms[*id])) return;
    Rect range = rooms[*id].has_range() ? rooms[*id].range : rooms[*id].assigned;
    for (size_t i = 0; i < range.len(); ++i) {
      Coord c = *range.nth(i);
      Cell* cell = field.try_get(c);
      if (!cell) throw GameError{ORC_ERR_SETTING, "in Floor::with_current_room"};
      mark(*cell, range.is_edge(c));
    }
  }
  void enters_room(Coord cd) {  // floor.rs:231-247
    with_current_room(
        cd,
        [](Room& room) {
          if (room.is_visited) return false;
          room.is_visited = true;
          return room.kind == KNormal && !room.is_dark;
        },
        [](Cell& c, bool) {
          c.attr |= HAS_DRAWN;
          c.visible(true);
        });
  }
  void leaves_room(Coord cd) {  // floor.rs:250-261
    with_current_room(
        cd, [](Room& room) { return room.is_visited && room.is_dark; },
        [](Cell& c, bool is_edge) {
          if (!is_edge) c.visible(false);
        });
  }

  // floor.rs:264-295; `activate_area` is passed in because EnemyHandler lives outside
  void player_in(Coord cd, bool init, const std::function<void(const Rect&)>& activate_area) {
    if (init || doors.count(cd)) {
      enters_room(cd);
      auto id = cd_to_room_id(cd);
      if (id) activate_area(rooms[*id].assigned);
    }
    Cell* cell = field.try_get(cd);
    if (!cell) throw GameError{ORC_ERR_SETTING, "Floor::player_in Cannot move"};
    cell->attr |= IS_VISITED;
    set_obj(cd, true);
    for (int d = 0; d < 9; ++d) {
      Cell* c = field.try_get(step(cd, d));
      if (!c) continue;
      if (!is_diag(d) || c->surface != SPassage) c->approached();
    }
  }
  void player_out(Coord cd) {  // floor.rs:298-312
    if (doors.count(cd)) leaves_room(cd);
    remove_obj(cd, true);
    for (int d = 0; d < 9; ++d) {
      Cell* c = field.try_get(step(cd, d));
      if (!c) continue;
      if (c->surface == SFloor) c->left();
    }
  }

  // floor.rs:349-370 ; returns number of SecretDoor messages
  int search(Coord cd, Rng& rng, const orc_params& cfg) {
    int secret = 0;
    const uint32_t probinc = 0;
    for (int d = 0; d < 8; ++d) {
      Cell* cell = field.try_get(step(cd, d));
      if (!cell) continue;
      if (cell->is_hidden() && rng.does_happen(probinc + cfg.passage_unlock_rate_inv)) {
        cell->unlock();
        cell->surface = SPassage;
      }
      if (cell->is_locked() && rng.does_happen(probinc + cfg.door_unlock_rate_inv)) {
        cell->unlock();
        cell->surface = SDoor;
        ++secret;
      }
    }
    return secret;
  }

  std::vector<uint8_t> history_map() const {  // floor.rs:372-379
    std::vector<uint8_t> h(field.cells.size());
    for (size_t i = 0; i < h.size(); ++i) h[i] = (field.cells[i].attr & IS_VISITED) ? 1 : 0;
    return h;
  }

  bool in_same_room(Coord a, Coord b) const {  // floor.rs:381-393
    auto id = cd_to_room_id(a);
    if (!id) return false;
    if (cd_to_room_id(b) != id) return false;
    const Room& r = rooms[*id];
    if (!r.has_range()) return true;
    return r.range.contains(a) == r.range.contains(b);
  }

  std::vector<uint32_t> make_dist_map(Coord from, bool is_enemy) const {  // floor.rs:395-416
    const uint32_t inf = UINT32_MAX;
    int w = field.w, h = field.h;
    std::vector<uint32_t> dist((size_t)w * h, inf);
    auto in = [&](Coord c) { return c.x >= 0 && c.y >= 0 && c.x < w && c.y < h; };
    if (!in(from)) panic("make_dist_map: from out of range");
    std::deque<Coord> q;
    dist[(size_t)from.y * w + from.x] = 0;
    q.push_back(from);
    while (!q.empty()) {
      Coord cur = q.front();
      q.pop_front();
      for (int d = 0; d < 8; ++d) {
        Coord nxt = step(cur, d);
        uint32_t cdist = dist[(size_t)cur.y * w + cur.x];
        if (!in(nxt)) continue;
        uint32_t& nd = dist[(size_t)nxt.y * w + nxt.x];
        if (nd != inf || can_move_impl(cur, d, is_enemy) != std::optional<bool>(true)) continue;
        q.push_back(nxt);
        nd = cdist + 1;
      }
    }
    return dist;
  }
};

// ------------------------------------------------------------------ the environment
enum : uint32_t {  // python/src/flags.rs:9-17
  MSG_HIT_FROM = 1, MSG_HIT_TO = 2, MSG_MISS_TO = 4, MSG_MISS_FROM = 8, MSG_KILLED = 16,
  MSG_SECRET_DOOR = 32, MSG_NO_DOWNSTAIR = 64
};
enum ReactKind { R_REDRAW, R_STATUS, R_DEAD, R_NOTIFY };
struct Reaction {
  int kind;
  uint32_t msg;  // message flag bit (0 for messages the bridge ignores)
};

struct Status {  // player.rs:389-399 in to_vec order
  uint32_t dungeon_level = 0, gold = 0, hp_cur = 0, hp_max = 0, str_cur = 0, str_max = 0, defense = 0,
           player_level = 0, exp = 0, hunger = 0;
  void to_vec(uint32_t* v) const {
    v[0] = dungeon_level; v[1] = gold; v[2] = hp_cur; v[3] = hp_max; v[4] = str_cur; v[5] = str_max;
    v[6] = defense; v[7] = player_level; v[8] = exp; v[9] = hunger;
  }
};

struct Env {
  orc_params cfg;
  int64_t max_steps;
  uint64_t seed_lo = 0, seed_hi = 0;
  std::string last_error;

  // --- RunTime (core/src/lib.rs:232-242)
  Rng rng_dungeon, rng_item, rng_enemy;
  uint32_t level = 0, max_level = 0;
  Floor floor;
  std::vector<std::vector<uint8_t>> past_history;  // history_map() of every left floor
  std::deque<std::pair<std::vector<uint32_t>, Coord>> dist_cache;  // rogue/mod.rs:492-518
  EnemyMap placed, active;
  // Player (player.rs:93-104,267-280)
  Coord ppos;
  int64_t hp = 0, hp_max = 0;
  uint32_t exp = 0;
  int64_t plevel = 1;
  uint32_t food_left = 0, quiet = 0;
  uint32_t pack_gold = 0;
  bool pack_has_gold_stack = false;
  bool ui_dead = false;

  // --- GameStateImpl / PlayerState (python/src/state_impls.rs:10-15, lib.rs:31-38)
  std::vector<uint8_t> screen, history;
  Status status;
  uint32_t message = 0;
  bool is_terminal = false;
  // what the last step RETURNED: ThreadConductor::step flags the copy it returns after an auto-reset
  // (thread_impls.rs:69-79); the worker's own state (is_terminal above, Instruction::State) is the fresh game's
  bool returned_terminal = false;
  int64_t steps = 0;
  int error = 0;

  explicit Env(const orc_params& p, int64_t ms) : cfg(p), max_steps(ms) {
    screen.assign((size_t)p.width * p.height, ' ');
    history.assign((size_t)p.width * p.height, 0);
  }

  Path path(Coord c) const { return Path{(int)level, c.x, c.y}; }

  // ------------------------------------------------ GameConfig::build core/src/lib.rs:193-228
  void build() {
    rng_item.seed(seed_lo, seed_hi);     // ItemHandler::new item/mod.rs:380-394
    rng_enemy.seed(seed_lo, seed_hi);    // enemies::Config::build enemies.rs:33-47
    rng_dungeon.seed(seed_lo, seed_hi);  // rogue::Dungeon::new rogue/mod.rs:417
    level = 0;
    max_level = cfg.amulet_level;
    floor = Floor();
    past_history.clear();
    dist_cache.clear();
    placed.clear();
    active.clear();
    new_level_(true);
    // Player::build player.rs:78-90 / StatusInner::from_config :283-293
    hp = hp_max = cfg.init_hp;
    exp = 0;
    plevel = 1;
    food_left = cfg.hunger_time;
    quiet = 0;
    ui_dead = false;
    // Player::init_items -> ItemHandler::init_player_items item/mod.rs:412-422: one
    // u32 draw per weapon preset on the ITEM stream (weapon.rs:159)
    for (uint32_t i = 0; i < cfg.n_init_draws; ++i) rng_item.range32(cfg.init_draw_lo[i], cfg.init_draw_hi[i]);
    pack_gold = cfg.init_gold;
    // actions::new_level(is_init = true) actions.rs:121-138
    auto cd = floor.select_cell(rng_dungeon, true);
    if (!cd) throw GameError{ORC_ERR_SETTING, "action::new_level No space for player!"};
    ppos = *cd;
    floor.player_in(ppos, true, [&](const Rect& area) { activate_area(area); });
  }

  uint32_t lev_add() const { return cfg.amulet_level < level ? level - cfg.amulet_level : 0; }  // rogue/mod.rs:483-489

  // rogue/mod.rs:434-481
  void new_level_(bool is_initial) {
    level += 1;
    if (level > max_level) max_level = level;
    Floor f = Floor::gen_floor(level, cfg, rng_dungeon);
    bool set_gold = true;  // !game_info.is_cleared || ... ; is_cleared is never set (lib.rs:417-423)
    // Floor::setup_items floor.rs:132-153
    if (set_gold) {
      for (Room& room : f.rooms) {
        auto cd = room.select_cell(rng_dungeon, false);
        if (!cd) continue;
        // gold::Config::gen item/gold.rs:18-24 on the ITEM stream
        if (!rng_item.does_happen(cfg.gold_rate_inv)) continue;
        uint32_t num = rng_item.range32(0, cfg.gold_base + cfg.gold_per_level * level) + cfg.gold_minimum;
        room.fill_cell(*cd, false);
        room.has_gold = true;
        f.items[*cd] = Item{num};
      }
    }
    {  // Floor::setup_stair floor.rs:156-167
      auto cd = f.select_cell(rng_dungeon, false);
      if (!cd) throw GameError{ORC_ERR_SETTING, "[setup stair] no empty cell!"};
      Cell* cell = f.field.try_get(*cd);
      if (!cell) throw GameError{ORC_ERR_SETTING, "[setup stair] select_cell returned invalid coord"};
      cell->surface = SStair;
      f.set_obj(*cd, false);
    }
    if (!is_initial) {  // EnemyHandler::remove_enemies enemies.rs:362-365
      placed.clear();
      active.clear();
    }
    // Floor::place_enemies floor.rs:106-130
    if (cfg.n_enemies != 0) {
      uint32_t mn = level >= 4 ? level - 4 : 0;
      uint32_t mx = level + 6;
      for (Room& room : f.rooms) {
        auto cd = room.select_cell(rng_dungeon, true);
        if (!cd) continue;
        auto en = gen_enemy(mn, mx, (int64_t)lev_add(), room.has_gold);
        if (en) {
          placed[Path{(int)level, cd->x, cd->y}] = en;
          room.fill_cell(*cd, true);
        }
      }
    }
    if (!cfg.hide_dungeon) {  // rogue/mod.rs:465-475
      for (int y = 1; y < cfg.height - 1; ++y)
        for (int x = 0; x < cfg.width; ++x) f.field.cells[(size_t)y * cfg.width + x].visible(true);
    }
    std::swap(floor, f);
    if (!is_initial) past_history.push_back(f.history_map());
  }

  // enemies.rs:265-320
  std::shared_ptr<Enemy> gen_enemy(uint32_t rmin, uint32_t rmax, int64_t lev_add_, bool has_gold) {
    uint32_t appear = has_gold ? cfg.appear_rate_gold : cfg.appear_rate_nogold;
    if (!rng_enemy.parcent(appear)) return nullptr;
    size_t len = cfg.n_enemies;
    size_t idx = (size_t)rng_enemy.range32(rmin, rmax);
    if (idx > len) {
      size_t r = std::min<size_t>(len, 5);
      idx = rng_enemy.range_usize(len - r, len);
    }
    if (idx >= len) return nullptr;
    const orc_enemy_kind& st = cfg.enemies[idx];
    int64_t lvl = (int64_t)st.level + lev_add_;
    int64_t hpv = 0;
    for (int i = 0; i < 8; ++i) hpv += rng_enemy.range_i64(1, lvl + 1);  // Dice::new(8, level).exec::<i64>
    auto e = std::make_shared<Enemy>();
    e->kind = (int)idx;
    e->attr = st.attr;
    e->defense = st.defense - (int32_t)lev_add_;
    int64_t base = (lvl == 1) ? hpv / 8 : hpv / 6;  // exp_add enemies.rs:275-285
    uint32_t add = (10 <= lvl) ? (uint32_t)base * 20u : (uint32_t)base * 4u;
    e->exp = st.exp + (uint32_t)(lev_add_ * 10) + add;
    e->hp = e->max_hp = hpv;
    e->level = lvl;
    e->running = false;
    return e;
  }

  std::shared_ptr<Enemy> get_enemy(const Path& p) const {  // enemies.rs:330-341
    auto it = placed.find(p);
    if (it != placed.end()) return it->second;
    auto jt = active.find(p);
    if (jt != active.end()) return jt->second;
    return nullptr;
  }
  bool activate(const Path& p) {  // enemies.rs:356-361
    auto it = placed.find(p);
    if (it == placed.end()) return false;
    auto e = it->second;
    placed.erase(it);
    e->running = true;
    active[p] = e;
    return true;
  }
  void activate_area(const Rect& area) {  // enemies.rs:342-355
    std::vector<Path> rm;
    for (auto& kv : placed)
      if (area.contains(Coord{kv.first[1], kv.first[2]}) && kv.second->is_mean()) rm.push_back(kv.first);
    for (auto& p : rm) activate(p);
  }

  // ---------------------------------------------------------------- DistCache + monster moves
  const std::vector<uint32_t>& cached_dist_map(Coord cd) {  // rogue/mod.rs:504-517
    for (auto& e : dist_cache)
      if (e.second == cd) return e.first;
    auto m = floor.make_dist_map(cd, true);
    size_t len = dist_cache.size();
    dist_cache.emplace_back(std::move(m), cd);
    if (len > 8) {
      dist_cache.pop_front();
      return dist_cache[len - 1].first;
    }
    return dist_cache[len].first;
  }

  MoveResult move_enemy(const Path& current, const Path& dist, const std::function<bool(const Path&)>& skip) {  // :339-375
    if (current[0] != dist[0]) return {CantMove, {}};
    Coord cur{current[1], current[2]};
    const std::vector<uint32_t>& dm = cached_dist_map(Coord{dist[1], dist[2]});
    int w = cfg.width, h = cfg.height;
    bool have = false;
    uint32_t best = 0;
    Coord best_cd;
    for (int d = 0; d < 9; ++d) {
      Coord next = step(cur, d);
      if (skip(Path{current[0], next.x, next.y})) continue;
      if (next.x < 0 || next.y < 0 || next.x >= w || next.y >= h) panic("dist_map.get_p out of range");  // :361
      uint32_t nd = dm[(size_t)next.y * w + next.x];
      if (nd == 0 && floor.can_move_enemy(cur, d)) return {Reach, {}};
      if (nd != UINT32_MAX && nd > 0) {
        if (!have || nd < best) {  // stable sort_by_key + [0]: first minimum in push order
          have = true;
          best = nd;
          best_cd = next;
        }
      }
    }
    if (!have) return {CantMove, {}};
    return {CanMove, Path{current[0], best_cd.x, best_cd.y}};
  }

  MoveResult move_enemy_randomly(const Path& enemy_pos, const Path& player_pos,
                                 const std::function<bool(const Path&)>& skip) {  // :376-397
    Coord cur{enemy_pos[1], enemy_pos[2]};
    size_t idx = rng_dungeon.range_usize(0, 8);
    int d = (int)idx;
    Coord next = step(cur, d);
    Path np{enemy_pos[0], next.x, next.y};
    if (skip(np) || !floor.can_move_enemy(cur, d)) return {CantMove, {}};
    if (np == player_pos) return {Reach, {}};
    return {CanMove, np};
  }

  std::vector<std::shared_ptr<Enemy>> move_actives(const Path& player_pos) {  // enemies.rs:366-424
    std::vector<std::shared_ptr<Enemy>> out;
    EnemyMap moving;
    std::swap(moving, active);
    for (auto& kv : moving) {
      const Path& p = kv.first;
      auto& enemy = kv.second;
      auto skip = [&](const Path& q) { return active.count(q) || placed.count(q); };
      // gold_pos is always None (actions.rs:88)
      MoveResult res;
      bool randomly;
      if (rng_enemy.does_happen(2) && enemy->is_random())
        randomly = true;
      else
        randomly = (!rng_enemy.does_happen(5) && enemy->is_confused());
      if (randomly)
        res = move_enemy_randomly(p, player_pos, skip);
      else
        res = move_enemy(p, player_pos, skip);
      Path next = p;
      if (res.kind == Reach)
        out.push_back(enemy);
      else if (res.kind == CanMove)
        next = res.to;
      active[next] = enemy;  // insert: overwrites a not-yet-moved monster standing there
    }
    return out;
  }

  // ------------------------------------------------------------------ fight.rs
  static uint32_t truncate_parcent(int64_t v) { return (uint32_t)std::min<int64_t>(100, std::max<int64_t>(0, v)); }
  static int64_t hit_prob_plus(int64_t st) {
    static const int64_t D[32] = {-7, -6, -5, -4, -3, -2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                  0,  1,  1,  1,  1,  2,  2,  2, 2, 2, 2, 2, 2, 2, 2, 3};
    if (st <= 0 || st > 32) return 0;
    return D[st - 1];
  }
  static int64_t damage_plus(int64_t st) {
    static const int64_t D[32] = {-7, -6, -5, -4, -3, -2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                  1,  1,  2,  3,  3,  4,  5,  5, 5, 5, 5, 5, 5, 5, 5, 6};
    if (st <= 0 || st > 32) return 0;
    return D[st - 1];
  }
  static constexpr int64_t PLAYER_STR = 16;  // player.rs:286
  static constexpr int64_t ENEMY_STR = 10;   // enemies.rs:173

  int64_t dice_random(int times, int64_t mx) { return dice_roll(rng_enemy, times, mx); }
  // fight.rs:52-72 ; nullopt = miss
  std::optional<int64_t> roll(const int32_t* times, const int32_t* maxs, int n, uint32_t rate, int64_t dam_plus) {
    bool did_hit = false;
    int64_t sum = 0;
    for (int i = 0; i < n; ++i) {
      if (!rng_enemy.parcent(rate)) continue;
      did_hit = true;
      sum += dice_random(times[i], maxs[i]) + dam_plus;
    }
    if (did_hit) return sum;
    return std::nullopt;
  }
  std::optional<int64_t> fight_player_attack(const Enemy& e) {  // fight.rs:6-39
    int64_t str_p = hit_prob_plus(PLAYER_STR) + (e.running ? 0 : 4) + cfg.weapon_hit_plus;
    uint32_t rate = truncate_parcent(((int64_t)plevel + e.defense + str_p + 1) * 5);
    int64_t dam_plus = cfg.weapon_dam_plus + damage_plus(PLAYER_STR);
    int32_t t = cfg.weapon_times, m = cfg.weapon_max;
    return roll(&t, &m, 1, rate, dam_plus);
  }
  std::optional<int64_t> fight_enemy_attack(const Enemy& e) {  // fight.rs:41-50
    uint32_t rate = truncate_parcent((e.level + cfg.armor_def + hit_prob_plus(ENEMY_STR) + 1) * 5);
    int64_t dam_plus = damage_plus(ENEMY_STR) + damage_plus(PLAYER_STR);
    const orc_enemy_kind& k = cfg.enemies[e.kind];
    return roll(k.dice_times, k.dice_max, (int)k.n_dice, rate, dam_plus);
  }

  // ------------------------------------------------------------------ player.rs
  bool player_get_damage(int64_t dmg) {  // :177-184 ; true = Death
    hp = std::max<int64_t>(hp - dmg, 0);
    return hp == 0;
  }
  bool level_up(uint32_t e) {  // :185-197, Leveling::check_level :346-352
    exp += e;
    size_t cur = (size_t)(plevel - 1);
    size_t diff = 0;
    if (cur < cfg.n_exps) {
      bool found = false;
      for (size_t i = cur; i < cfg.n_exps; ++i)
        if (exp < cfg.exps[i]) {
          diff = i - cur;
          found = true;
          break;
        }
      if (!found) panic("check_level: unwrap on None");
    }
    if (diff > 0) {
      plevel += (int64_t)diff;
      int64_t add = 0;
      for (size_t i = 0; i < diff; ++i) add += rng_enemy.range_i64(1, 11);  // Dice::new(diff, 10).exec::<i64>
      hp_max += add;
      hp += add;
      return true;
    }
    return false;
  }
  bool heal() {  // :221-240
    quiet += 1;
    int64_t q = quiet;
    int64_t lv = plevel;
    int64_t amount;
    if (lv < 8)
      amount = std::max<int64_t>(std::min<int64_t>(q + (lv << 1) - 20, 1), 0);
    else if (q >= 3)
      amount = rng_enemy.range_i64(1, lv - 6);
    else
      amount = 0;
    if (amount > 0) {
      hp += amount;
      if (hp > hp_max) hp = hp_max;
      quiet = 0;
      return true;
    }
    return false;
  }
  // Player::turn_passed :163-176 folded with actions::after_turn's event handling (actions.rs:73-78)
  void turn_passed(std::vector<Reaction>& res) {
    food_left -= 1;  // u32, wraps (release build)
    if (food_left == 0) return;  // [Dead] is ignored by the caller
    uint32_t hunger = cfg.hunger_time / 10;
    bool hungry = (food_left == hunger || food_left == hunger * 2);
    bool healed = heal();
    if (hungry) res.push_back({R_STATUS, 0});
    if (healed) res.push_back({R_STATUS, 0});
  }
  Status player_status() const {  // core/src/lib.rs:345-356 + Player::fill_status player.rs:107-118
    Status s;
    s.hp_cur = (uint32_t)hp;
    s.hp_max = (uint32_t)hp_max;
    s.str_cur = s.str_max = (uint32_t)PLAYER_STR;
    s.exp = exp;
    s.player_level = (uint32_t)plevel;
    uint32_t hunger = cfg.hunger_time / 10;
    s.hunger = food_left <= hunger ? 2u : (food_left <= hunger * 2 ? 1u : 0u);
    s.defense = 0;
    s.gold = pack_gold;
    s.dungeon_level = level;
    return s;
  }

  // ------------------------------------------------------------------ actions.rs
  bool move_active_enemies(std::vector<Reaction>& res) {  // :82-119 ; true => Some(Grave)
    auto attacks = move_actives(path(ppos));
    if (!attacks.empty()) quiet = 0;  // player.buttle()
    bool did_hit = false;
    for (auto& en : attacks) {
      auto dmg = fight_enemy_attack(*en);
      if (dmg) {
        res.push_back({R_NOTIFY, MSG_HIT_FROM});
        did_hit = true;
        if (player_get_damage(*dmg)) {
          res.push_back({R_DEAD, 0});
          return true;
        }
      } else {
        res.push_back({R_NOTIFY, MSG_MISS_FROM});
      }
    }
    if (did_hit) res.push_back({R_STATUS, 0});
    return false;
  }
  bool after_turn(std::vector<Reaction>& res) {  // :67-80
    turn_passed(res);
    return move_active_enemies(res);
  }

  std::vector<Reaction> player_attack(std::shared_ptr<Enemy> enemy, const Path& place) {  // :140-166
    std::vector<Reaction> res;
    quiet = 0;
    activate(place);
    auto dmg = fight_player_attack(*enemy);
    if (dmg) {
      res.push_back({R_NOTIFY, MSG_HIT_TO});
      // Enemy::get_damage enemies.rs:205-213
      if (enemy->hp <= *dmg) {
        placed.erase(place);
        active.erase(place);
        if (level_up(enemy->exp)) res.push_back({R_STATUS, 0});
        res.push_back({R_NOTIFY, MSG_KILLED});
        res.push_back({R_REDRAW, 0});
      } else {
        enemy->hp = *dmg - enemy->hp;
      }
    } else {
      res.push_back({R_NOTIFY, MSG_MISS_TO});
    }
    return res;
  }

  std::pair<std::vector<Reaction>, bool> move_player(int d) {  // :168-195
    auto np = floor.can_move_player(ppos, d);
    if (!np) return {{Reaction{R_NOTIFY, 0}}, true};  // CantMove: not a message flag
    Path npath = path(*np);
    if (auto en = get_enemy(npath)) return {player_attack(en, npath), true};
    // rogue::Dungeon::move_player rogue/mod.rs:237-258
    floor.player_out(ppos);
    Coord cd = step(ppos, d);
    floor.player_in(cd, false, [&](const Rect& area) { activate_area(area); });
    ppos = cd;
    bool done = false;
    std::vector<Reaction> res{Reaction{R_REDRAW, 0}};
    // get_item :206-231
    auto it = floor.items.find(ppos);
    if (it != floor.items.end() && cfg.pack_accepts_gold) {
      pack_gold += it->second.amount;  // MergeEntry / InsertEntry
      if (floor.remove_obj(ppos, false)) floor.items.erase(it);  // rogue/mod.rs:311-320
      res.push_back({R_NOTIFY, 0});  // GotItem: not a message flag
      res.push_back({R_STATUS, 0});
      done = true;
    }
    return {res, done};
  }

  // process_action :16-65 ; key already mapped. Returns (ui_dead, reactions)
  std::pair<bool, std::vector<Reaction>> process_action(int act, int d) {
    std::vector<Reaction> out;
    bool ui = false;
    switch (act) {
      case 3: {  // DownStair
        const Cell* c = floor.field.try_get(ppos);
        if (c && c->surface == SStair) {
          // actions::new_level(is_init=false)
          new_level_(false);
          auto cd = floor.select_cell(rng_dungeon, true);
          if (!cd) throw GameError{ORC_ERR_SETTING, "action::new_level No space for player!"};
          ppos = *cd;
          floor.player_in(ppos, true, [&](const Rect& area) { activate_area(area); });
          out.push_back({R_REDRAW, 0});
          out.push_back({R_STATUS, 0});
        } else {
          out.push_back({R_NOTIFY, MSG_NO_DOWNSTAIR});
        }
        ui = after_turn(out);
        break;
      }
      case 0: {  // Move
        auto r = move_player(d);
        out.insert(out.end(), r.first.begin(), r.first.end());
        ui = after_turn(out);
        break;
      }
      case 1: {  // MoveUntil
        for (;;) {
          auto r = move_player(d);
          const Cell* c = floor.field.try_get(ppos);
          char tile = c ? (char)c->tile() : ' ';
          if (r.second || (tile != '.' && tile != '#')) {
            out.insert(out.end(), r.first.begin(), r.first.end());
            break;
          } else if (out.empty()) {
            out.insert(out.end(), r.first.begin(), r.first.end());
          }
          ui = after_turn(out);
        }
        break;
      }
      case 2: {  // Search
        int secret = floor.search(ppos, rng_dungeon, cfg);
        for (int i = 0; i < secret; ++i) out.push_back({R_NOTIFY, MSG_SECRET_DOOR});
        out.push_back({R_REDRAW, 0});
        ui = after_turn(out);
        break;
      }
      default:  // NoOp
        return {false, out};
    }
    return {ui, out};
  }

  // ---------------------------------------------------- PlayerState (python/src/lib.rs:41-68)
  void draw_map() {
    // RunTime::history -> Dungeon::get_history rogue/mod.rs:329-338 with the DISPLAYED level
    uint32_t lvl = status.dungeon_level;
    if (lvl == level)
      history = floor.history_map();
    else if (lvl >= 1 && (size_t)(lvl - 1) < past_history.size())
      history = past_history[lvl - 1];
    else
      panic("history unwrap on None");
    int w = cfg.width, h = cfg.height;
    // Dungeon::draw rogue/mod.rs:278-290
    for (int y = 1; y < h - 1; ++y)
      for (int x = 0; x < w; ++x) screen[(size_t)y * w + x] = floor.field.cells[(size_t)y * w + x].tile();
    // draw_ranges + overlay core/src/lib.rs:270-284
    for (int y = 1; y < h - 1; ++y)
      for (int x = 0; x < w; ++x) {
        const Cell& c = floor.field.cells[(size_t)y * w + x];
        if (!c.is_obj_visible()) continue;
        Coord cd{x, y};
        if (ppos == cd) {
          screen[(size_t)y * w + x] = '@';
          continue;
        }
        if (floor.items.count(cd)) {
          screen[(size_t)y * w + x] = '*';
          continue;
        }
        if (auto en = get_enemy(path(cd))) {
          // Dungeon::draw_enemy rogue/mod.rs:398-404
          int ddx = ppos.x - cd.x, ddy = ppos.y - cd.y;
          bool adjacent = ddx * ddx + ddy * ddy <= 2;
          if (adjacent || floor.in_same_room(ppos, cd)) screen[(size_t)y * w + x] = (uint8_t)cfg.enemies[en->kind].tile;
        }
      }
  }
  void state_reset() {  // PlayerState::reset lib.rs:52-58
    status = player_status();
    draw_map();
    message = 0;
    is_terminal = false;
  }

  // ---------------------------------------------------- GameStateImpl (state_impls.rs)
  void reset() {
    build();
    state_reset();
    steps = 0;
  }

  static bool map_key(uint8_t key, int* act, int* d) {  // KeyMap::ai input.rs:74-99
    switch (key) {
      case 'l': *act = 0; *d = Right; return true;
      case 'k': *act = 0; *d = Up; return true;
      case 'j': *act = 0; *d = Down; return true;
      case 'h': *act = 0; *d = Left; return true;
      case 'u': *act = 0; *d = RightUp; return true;
      case 'y': *act = 0; *d = LeftUp; return true;
      case 'n': *act = 0; *d = RightDown; return true;
      case 'b': *act = 0; *d = LeftDown; return true;
      case '.': *act = 4; *d = Stay; return true;
      case 'L': *act = 1; *d = Right; return true;
      case 'K': *act = 1; *d = Up; return true;
      case 'J': *act = 1; *d = Down; return true;
      case 'H': *act = 1; *d = Left; return true;
      case 'U': *act = 1; *d = RightUp; return true;
      case 'Y': *act = 1; *d = LeftUp; return true;
      case 'N': *act = 1; *d = RightDown; return true;
      case 'B': *act = 1; *d = LeftDown; return true;
      case 's': *act = 2; *d = Stay; return true;
      case '>': *act = 3; *d = Stay; return true;
      default: return false;
    }
  }

  void react(uint8_t key) {  // state_impls.rs:51-79
    if (steps > max_steps) return;
    int act, d;
    if (!map_key(key, &act, &d)) throw GameError{ORC_ERR_INVALID_INPUT, "Invliad input key"};
    if (ui_dead) throw GameError{ORC_ERR_IGNORED_INPUT, "Ignored input code"};  // core/src/lib.rs:314
    auto pr = process_action(act, d);
    if (pr.first) ui_dead = true;  // `if let Some(next_ui)` core/src/lib.rs:317-319
    message = 0;
    bool dead = false;
    for (const Reaction& r : pr.second) {
      switch (r.kind) {
        case R_REDRAW: draw_map(); break;
        case R_STATUS: status = player_status(); break;
        case R_DEAD: dead = true; break;
        case R_NOTIFY: message |= r.msg; break;
      }
    }
    steps += 1;
    is_terminal = dead || steps >= max_steps;
  }
};

uint64_t fnv(uint64_t h, const void* p, size_t n) {
  const uint8_t* b = (const uint8_t*)p;
  for (size_t i = 0; i < n; ++i) {
    h ^= b[i];
    h *= 0x100000001b3ull;
  }
  return h;
}

template <class F>
int guarded(Env* e, F&& f) {
  if (e->error == ORC_ERR_PANIC || e->error == ORC_ERR_SETTING) return e->error;  // sticky
  try {
    f();
    return ORC_OK;
  } catch (const Panic& p) {
    e->error = ORC_ERR_PANIC;
    e->last_error = "panic: " + p.msg;
    return e->error;
  } catch (const GameError& g) {
    if (g.code == ORC_ERR_SETTING) e->error = g.code;
    e->last_error = g.msg;
    return g.code;
  }
}

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

}  // namespace

extern "C" {

void* orc_create(const orc_params* p, int64_t max_steps) { return new Env(*p, max_steps); }
void orc_destroy(void* env) { delete (Env*)env; }
void orc_set_seed(void* env, uint64_t lo, uint64_t hi) {
  Env* e = (Env*)env;
  e->seed_lo = lo;
  e->seed_hi = hi;
}
int orc_reset(void* env) {
  Env* e = (Env*)env;
  e->error = 0;
  e->returned_terminal = false;
  return guarded(e, [&] { e->reset(); });
}
int orc_react(void* env, uint8_t key) {
  Env* e = (Env*)env;
  int rc = guarded(e, [&] { e->react(key); });
  if (rc == ORC_OK) e->returned_terminal = e->is_terminal;
  return rc;
}
int orc_step_auto(void* env, uint8_t key) {
  Env* e = (Env*)env;
  int rc = guarded(e, [&] { e->react(key); });
  if (rc != ORC_OK) return rc;
  e->returned_terminal = e->is_terminal;
  if (e->is_terminal) rc = guarded(e, [&] { e->reset(); });  // the fresh state itself is not terminal
  return rc;
}
const char* orc_last_error(void* env) { return ((Env*)env)->last_error.c_str(); }

void orc_get_obs(void* env, uint8_t* screen, uint8_t* history, uint32_t* status10, uint32_t* message,
                 int32_t* is_terminal) {
  Env* e = (Env*)env;
  if (screen) memcpy(screen, e->screen.data(), e->screen.size());
  if (history) memcpy(history, e->history.data(), e->history.size());
  if (status10) e->status.to_vec(status10);
  if (message) *message = e->message;
  if (is_terminal) *is_terminal = e->returned_terminal;
}

void orc_get_scalars(void* env, orc_scalars* o) {
  Env* e = (Env*)env;
  memset(o, 0, sizeof(*o));
  o->level = (int32_t)e->level;
  o->px = e->ppos.x;
  o->py = e->ppos.y;
  o->hp = (int32_t)e->hp;
  o->hp_max = (int32_t)e->hp_max;
  o->exp = e->exp;
  o->plevel = (int32_t)e->plevel;
  o->food_left = e->food_left;
  o->quiet = e->quiet;
  o->gold = e->pack_gold;
  o->ui_dead = e->ui_dead;
  o->steps = (int32_t)e->steps;
  o->is_terminal = e->is_terminal;
  o->message = e->message;
  o->error = e->error;
  o->n_monsters = (int32_t)(e->placed.size() + e->active.size());
  o->n_items = (int32_t)e->floor.items.size();
  o->n_cache = (int32_t)e->dist_cache.size();
  e->status.to_vec(o->status);
  const Rng* r[3] = {&e->rng_dungeon, &e->rng_item, &e->rng_enemy};
  for (int i = 0; i < 3; ++i) {
    o->rng[i * 4 + 0] = r[i]->x;
    o->rng[i * 4 + 1] = r[i]->y;
    o->rng[i * 4 + 2] = r[i]->z;
    o->rng[i * 4 + 3] = r[i]->w;
  }
}

void orc_get_grid(void* env, uint8_t* surface, uint8_t* attr) {
  Env* e = (Env*)env;
  const Field& f = e->floor.field;
  for (size_t i = 0; i < f.cells.size(); ++i) {
    surface[i] = f.cells[i].surface;
    attr[i] = (uint8_t)(f.cells[i].attr & 63);
  }
  for (const Coord& c : e->floor.doors) attr[(size_t)c.y * f.w + c.x] |= 64;
}

void orc_get_entities(void* env, int32_t* monsters, int32_t* items) {
  Env* e = (Env*)env;
  memset(monsters, 0, sizeof(int32_t) * ORC_MAX_ROOMS * 8);
  memset(items, 0, sizeof(int32_t) * ORC_MAX_ROOMS * 3);
  std::map<Path, std::pair<std::shared_ptr<Enemy>, int>> all;
  for (auto& kv : e->placed) all[kv.first] = {kv.second, 0};
  for (auto& kv : e->active) all[kv.first] = {kv.second, 1};
  int i = 0;
  for (auto& kv : all) {
    if (i >= ORC_MAX_ROOMS) break;
    int32_t* m = monsters + i * 8;
    m[0] = kv.first[1];
    m[1] = kv.first[2];
    m[2] = kv.second.first->kind;
    m[3] = (int32_t)kv.second.first->hp;
    m[4] = kv.second.second;
    m[5] = (int32_t)kv.second.first->level;
    m[6] = kv.second.first->defense;
    m[7] = (int32_t)kv.second.first->exp;
    ++i;
  }
  std::vector<std::array<int32_t, 3>> its;
  for (auto& kv : e->floor.items) its.push_back({kv.first.x, kv.first.y, (int32_t)kv.second.amount});
  int w = e->cfg.width;
  std::sort(its.begin(), its.end(), [w](auto& a, auto& b) { return a[1] * w + a[0] < b[1] * w + b[0]; });
  for (size_t k = 0; k < its.size() && k < ORC_MAX_ROOMS; ++k)
    for (int j = 0; j < 3; ++j) items[k * 3 + j] = its[k][j];
}

void orc_get_dist_cache(void* env, int32_t* xy, uint16_t* maps) {
  Env* e = (Env*)env;
  size_t C = (size_t)e->cfg.width * e->cfg.height;
  for (int i = 0; i < ORC_DIST_CACHE; ++i) {
    xy[i * 2] = xy[i * 2 + 1] = -1;
  }
  for (size_t i = 0; i < e->dist_cache.size() && i < ORC_DIST_CACHE; ++i) {
    xy[i * 2] = e->dist_cache[i].second.x;
    xy[i * 2 + 1] = e->dist_cache[i].second.y;
    if (maps)
      for (size_t c = 0; c < C; ++c) {
        uint32_t v = e->dist_cache[i].first[c];
        maps[i * C + c] = v == UINT32_MAX ? 0xFFFF : (uint16_t)v;
      }
  }
}

void orc_get_rooms(void* env, int32_t* rooms) {
  Env* e = (Env*)env;
  memset(rooms, 0, sizeof(int32_t) * ORC_MAX_ROOMS * 8);
  for (size_t i = 0; i < e->floor.rooms.size() && i < ORC_MAX_ROOMS; ++i) {
    const Room& r = e->floor.rooms[i];
    int32_t* o = rooms + i * 8;
    o[0] = r.kind;
    o[1] = r.is_dark;
    o[2] = r.is_visited;
    o[3] = r.has_gold;
    if (r.kind == KEmpty) {
      o[4] = r.up_left.x;
      o[5] = r.up_left.y;
      o[6] = r.up_left.x;
      o[7] = r.up_left.y;
    } else {
      o[4] = r.range.x0;
      o[5] = r.range.y0;
      o[6] = r.range.x1;
      o[7] = r.range.y1;
    }
  }
}

void orc_get_draw_counts(void* env, uint64_t* c) {
  Env* e = (Env*)env;
  c[0] = e->rng_dungeon.draws;
  c[1] = e->rng_item.draws;
  c[2] = e->rng_enemy.draws;
}

static int sym_of(uint8_t t) {  // symbol.rs:17-40
  switch (t) {
    case ' ': return 0;
    case '@': return 1;
    case '#': return 2;
    case '.': return 3;
    case '-':
    case '|': return 4;
    case '%': return 5;
    case '+': return 6;
    case '^': return 7;
    case '!': return 8;
    case '?': return 9;
    case ']': return 10;
    case ')': return 11;
    case '/': return 12;
    case '*': return 13;
    case ':': return 14;
    case '=': return 15;
    case ',': return 16;
    default:
      if (t >= 'A' && t <= 'Z') return t - 'A' + 17;
      return -1;
  }
}

int orc_encode(void* env, int mode, uint32_t flag, int with_hist, float* out) {
  Env* e = (Env*)env;
  size_t C = (size_t)e->cfg.width * e->cfg.height;
  int symbols = (int)e->cfg.symbols;
  int nstat = __builtin_popcount(flag & 0x1FF);
  int base = mode == 0 ? 1 : symbols;
  int channels = base + nstat + (with_hist ? 1 : 0);
  std::fill(out, out + (size_t)channels * C, 0.0f);
  if (mode == 0) {  // gray_image_with_offset lib.rs:72-87
    for (size_t c = 0; c < C; ++c) {
      int s = sym_of(e->screen[c]);
      if (s < 0) return -1;
      out[c] = (float)s / (float)symbols;
    }
  } else {  // symbol_image_with_offset :88-104 -> construct_symbol_map(symbol_max = symbols - 1)
    int symbol_max = symbols - 1;
    for (int i = 0; i < symbol_max; ++i)
      for (size_t c = 0; c < C; ++c) {
        int s = sym_of(e->screen[c]);
        if (s < 0 || s >= symbol_max) return -1;  // InvalidTileError
        out[(size_t)i * C + c] = (s == i) ? 1.0f : 0.0f;
      }
  }
  // StatusFlagInner::copy_status flags.rs:87-115
  uint32_t st[10];
  e->status.to_vec(st);
  const int order[9] = {0, 2, 3, 4, 5, 6, 7, 8, 9};
  int off = base;
  for (int b = 0; b < 9; ++b)
    if (flag & (1u << b)) {
      float v = (float)(int32_t)st[order[b]];
      std::fill(out + (size_t)off * C, out + (size_t)(off + 1) * C, v);
      ++off;
    }
  if (with_hist) {  // copy_hist lib.rs:105-111
    for (size_t c = 0; c < C; ++c) out[(size_t)off * C + c] = e->history[c] ? 1.0f : 0.0f;
  }
  return channels;
}

int orc_test_move_enemy(void* env, int fx, int fy, int tx, int ty, int* nx, int* ny) {
  Env* e = (Env*)env;
  int kind = -1;
  guarded(e, [&] {
    MoveResult r = e->move_enemy(Path{(int)e->level, fx, fy}, Path{(int)e->level, tx, ty}, [](const Path&) { return false; });
    kind = r.kind;
    if (r.kind == CanMove) {
      *nx = r.to[1];
      *ny = r.to[2];
    }
  });
  return kind;
}

uint64_t orc_state_hash(void* env) {
  Env* e = (Env*)env;
  uint64_t h = 0xcbf29ce484222325ull;
  h = fnv(h, e->screen.data(), e->screen.size());
  uint32_t st[10];
  e->status.to_vec(st);
  h = fnv(h, st, sizeof(st));
  uint32_t r[12] = {e->rng_dungeon.x, e->rng_dungeon.y, e->rng_dungeon.z, e->rng_dungeon.w,
                    e->rng_item.x,    e->rng_item.y,    e->rng_item.z,    e->rng_item.w,
                    e->rng_enemy.x,   e->rng_enemy.y,   e->rng_enemy.z,   e->rng_enemy.w};
  h = fnv(h, r, sizeof(r));
  int32_t s[6] = {e->ppos.x, e->ppos.y, (int32_t)e->hp, (int32_t)e->level, (int32_t)e->steps, (int32_t)e->is_terminal};
  h = fnv(h, s, sizeof(s));
  return h;
}

double orc_batch_rollout(void** envs, int64_t n, int64_t first_env_id, int64_t t0, int64_t steps, int threads,
                         int /*unused*/, uint64_t* digest) {
  static const char KEYS[] = ".hjklnbuy>s";
  if (threads < 1) threads = 1;
  std::vector<uint64_t> part((size_t)threads, 0);
  auto work = [&](int tid) {
    int64_t lo = n * tid / threads, hi = n * (tid + 1) / threads;
    uint64_t acc = 0;
    for (int64_t i = lo; i < hi; ++i) {
      for (int64_t t = t0; t < t0 + steps; ++t) {
        uint64_t a = splitmix64((0x9E3779B97F4A7C15ull * (uint64_t)(t + 1)) ^ (uint64_t)(first_env_id + i)) % 11;
        orc_step_auto(envs[i], (uint8_t)KEYS[a]);
      }
      acc ^= orc_state_hash(envs[i]) * (uint64_t)(2 * (first_env_id + i) + 1);
    }
    part[(size_t)tid] = acc;
  };
  auto t_start = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  auto t_end = std::chrono::steady_clock::now();
  uint64_t d = 0;
  for (uint64_t v : part) d ^= v;
  if (digest) *digest = d;
  return std::chrono::duration<double>(t_end - t_start).count();
}

// ---- hooks for the reference's own unit-test known answers (fenwick.rs:262-305, passages.rs:272-296)
// A set over [0, cap) holding `members`; returns nth(k) or -1. ops (nullable): per member 1 = insert, 2 = remove,
// applied in order, with each call's bool result written back to ops_result.
int64_t orc_test_ordset(uint64_t cap, const uint64_t* members, const uint8_t* ops, uint8_t* ops_result, int64_t n,
                        int64_t k, uint64_t* len_out) {
  OrdSet s((size_t)cap);
  for (int64_t i = 0; i < n; ++i) {
    bool r = (ops && ops[i] == 2) ? s.remove((size_t)members[i]) : s.insert((size_t)members[i]);
    if (ops_result) ops_result[i] = r ? 1 : 0;
  }
  if (len_out) *len_out = s.len();
  auto v = s.nth((size_t)k);
  return v ? (int64_t)*v : -1;
}
// character/mod.rs:277-285 `test_dice`: `count` rolls of `times` d `mx` on a stream seeded with `seed`
void orc_test_dice(uint64_t seed, int times, int64_t mx, int64_t count, int64_t* out) {
  Rng r;
  r.seed(seed, 0);
  for (int64_t i = 0; i < count; ++i) out[i] = dice_roll(r, times, mx);
}
// floor.rs:490-505 `select_cell`: a floor of `level`, then select_cell(false) + set_obj until no cell is left or
// `tries` are used up. Returns how many cells were handed out, -1 if set_obj refused one of them.
int64_t orc_test_select_cell(const orc_params* p, uint32_t level, uint64_t seed, int64_t tries) {
  Rng r;
  r.seed(seed, 0);
  Floor f = Floor::gen_floor(level, *p, r);
  int64_t cnt = 0;
  for (int64_t i = 0; i < tries; ++i) {
    auto cd = f.select_cell(r, false);
    if (!cd) break;
    if (!f.set_obj(*cd, false)) return -1;
    ++cnt;
  }
  return cnt;
}
int orc_test_ordset_from_range_contains(uint64_t lo, uint64_t hi, uint64_t e) {
  return OrdSet::from_range((size_t)lo, (size_t)hi).contains((size_t)e) ? 1 : 0;
}
// passages::edges of the half-open rect x0..x1 x y0..y1; direction in enum-iterator order (Up 0, Down 1, Left 2,
// Right 3); writes (x, y) pairs, returns how many.
int orc_test_edges(int x0, int x1, int y0, int y1, int direction, int inclusive, int32_t* out_xy, int cap) {
  Rect r{x0, y0, x1, y1};
  auto v = edges(r, direction, inclusive != 0);
  int n = 0;
  for (Coord c : v) {
    if (n >= cap) break;
    out_xy[2 * n] = c.x;
    out_xy[2 * n + 1] = c.y;
    ++n;
  }
  return (int)v.size();
}

// ---- batch helpers for lock-step parity tests (no reference counterpart)
void orc_batch_step(void** envs, int64_t n, const uint8_t* keys, int auto_reset, int threads, int32_t* rc_out) {
  if (threads < 1) threads = 1;
  auto work = [&](int tid) {
    int64_t lo = n * tid / threads, hi = n * (tid + 1) / threads;
    for (int64_t i = lo; i < hi; ++i) {
      int rc = auto_reset ? orc_step_auto(envs[i], keys[i]) : orc_react(envs[i], keys[i]);
      if (rc_out) rc_out[i] = rc;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}
void orc_batch_reset(void** envs, int64_t n, int threads, int32_t* rc_out) {
  if (threads < 1) threads = 1;
  auto work = [&](int tid) {
    int64_t lo = n * tid / threads, hi = n * (tid + 1) / threads;
    for (int64_t i = lo; i < hi; ++i) {
      int rc = orc_reset(envs[i]);
      if (rc_out) rc_out[i] = rc;
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
}
void orc_batch_get_obs(void** envs, int64_t n, uint8_t* screen, uint8_t* history, uint32_t* status, uint32_t* message,
                       uint8_t* is_terminal) {
  for (int64_t i = 0; i < n; ++i) {
    Env* e = (Env*)envs[i];
    size_t C = e->screen.size();
    int32_t term = 0;
    uint32_t msg = 0;
    orc_get_obs(envs[i], screen ? screen + i * C : nullptr, history ? history + i * C : nullptr,
                status ? status + i * 10 : nullptr, &msg, &term);
    if (message) message[i] = msg;
    if (is_terminal) is_terminal[i] = (uint8_t)term;
  }
}
void orc_batch_hash(void** envs, int64_t n, uint64_t* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = orc_state_hash(envs[i]);
}

}  // extern "C"
