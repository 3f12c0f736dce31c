// rg_gen.inl - the floor generator (the functions that draw from the RNG streams). Included twice by
// rg_device.cuh, in namespace gen_inl (samplers inlined: lowest latency for the one or two warps of a
// descent inside a step) and in namespace gen_call (samplers as shared non-inlined functions: ~100 call
// sites share one copy in the instruction caches; +41 % floors/s when thousands of warps generate).
// The ...G names are macros set by the includer to the inlined or the called variant of Rng's samplers.

#ifndef RG_GEN_SECOND_HALF
// ------------------------------------------------------------------ floor generation
// maze::dig_maze / dig_impl maze.rs:38-89. The recursion is an explicit stack kept in this
// env's (not yet composed) screen slice; membership = A_MARK.
__device__ int dig_maze(Ctx& c, Rng& r, int x0, int y0, int x1, int y1) {
  RG_PLANES(c);
  uint16_t* stack = reinterpret_cast<uint16_t*>(c.g_screen);
  const int W = c.W;
  int sp = 0, n = 1;
  int cx = x0, cy = y0;
  A[cy * W + cx] |= A_MARK;
  const int max_sp = c.CP / 2;
  for (;;) {
    int pick = -1;
    uint32_t k = 0;
    // the four membership reads are issued together (one shared-memory latency instead of four on this
    // serial chain); an out-of-range neighbour reads the current cell, which is marked
    bool open[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int tx = cx + 2 * ddx(d), ty = cy + 2 * ddy(d);
      const bool inside = tx >= x0 && tx < x1 && ty >= y0 && ty < y1;
      open[d] = !(A[inside ? ty * W + tx : cy * W + cx] & A_MARK);
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      if (open[d]) {
        if (r.does_happenG(k + 1)) pick = d;  // reservoir pick: every candidate draws (maze.rs:64-75)
        ++k;
      }
    }
    if (pick < 0) {
      if (sp == 0) break;
      uint16_t v = stack[--sp];
      cx = v & 0xff;
      cy = v >> 8;
      continue;
    }
    for (int s = 1; s <= 2; ++s) {
      int mi = (cy + s * ddy(pick)) * W + cx + s * ddx(pick);
      if (!(A[mi] & A_MARK)) {
        A[mi] |= A_MARK;
        ++n;
      }
    }
    if (sp >= max_sp) {
      set_panic(c);
      break;
    }
    stack[sp++] = (uint16_t)(cx | (cy << 8));
    cx += 2 * ddx(pick);
    cy += 2 * ddy(pick);
  }
  return n;
}

// rooms::gen_rooms + make_room rooms.rs:165-269 (geometry and RNG only; tiles are laid later)
__device__ void gen_rooms(Ctx& c, Rng& r, uint32_t level) {
  RG_PLANES(c);
  const rg_params& P = *c.P;
  const int nrooms = c.nrooms;
  uint32_t empty_num = r.range32G(0, P.max_empty_rooms + 1);
  if (empty_num >= (uint32_t)nrooms) empty_num = (uint32_t)nrooms - 1;
  uint32_t rest = (nrooms >= 32) ? RG_FULL : ((1u << nrooms) - 1u), empty_mask = 0;
  for (uint32_t i = 0; i < empty_num; ++i) {  // RandomSelecter rng.rs:129-143
    uint32_t cnt = __popc(rest);
    if (!cnt) break;
    uint32_t n = (uint32_t)r.range64G(0, cnt);
    int b = nth_set_bit(rest, n);
    rest &= ~(1u << b);
    empty_mask |= 1u << b;
  }
  for (int i = 0; i < nrooms; ++i) {
    int ax0, ay0, ax1, ay1;
    room_area(c, i, ax0, ay0, ax1, ay1);
    int sx = ax1 - ax0, sy = ay1 - ay0;
    RoomD rm;
    rm.ncells = 0;
    if ((empty_mask >> i) & 1u) {
      int x = r.range_i32G(1, sx - 1) + ax0;
      int y = r.range_i32G(1, sy - 1) + ay0;
      rm.kind = K_EMPTY;
      rm.flags = RF_DARK;
      rm.x0 = rm.x1 = (uint8_t)x;
      rm.y0 = rm.y1 = (uint8_t)y;
    } else {
      bool dark = r.range32G(0, P.dark_level) < level;
      rm.flags = dark ? RF_DARK : 0;
      if (dark && r.does_happenG(P.maze_rate_inv)) {
        rm.kind = K_MAZE;
        rm.x0 = (uint8_t)ax0;
        rm.y0 = (uint8_t)ay0;
        rm.x1 = (uint8_t)(ax1 - 1);
        rm.y1 = (uint8_t)(ay1 - 1);
        rm.ncells = (uint16_t)dig_maze(c, r, ax0, ay0, ax1 - 1, ay1 - 1);
      } else {
        rm.kind = K_NORMAL;
        int w = r.range_i32G(P.min_room_x, sx);
        int h = r.range_i32G(P.min_room_y, sy);
        int lx = r.range_i32G(0, sx - w) + ax0;
        int ly = r.range_i32G(0, sy - h) + ay0;
        rm.x0 = (uint8_t)lx;
        rm.y0 = (uint8_t)ly;
        rm.x1 = (uint8_t)(lx + w);
        rm.y1 = (uint8_t)(ly + h);
      }
    }
    st->rooms[i] = rm;
  }
}

// Room::draw + gen_attr for room tiles: floor.rs:61-71,420-451, rooms.rs:58-82
__device__ void lay_rooms(Ctx& c, Rng& r, uint32_t level) {
  RG_PLANES(c);
  const rg_params& P = *c.P;
  const int W = c.W;
  for (int i = 0; i < c.nrooms; ++i) {
    RoomD rm = st->rooms[i];
    if (rm.kind == K_NORMAL) {  // PARALLEL: no RNG is consumed by wall / floor tiles
      int w = rm.x1 - rm.x0, n = w * (rm.y1 - rm.y0);
      uint8_t fl = (rm.flags & RF_DARK) ? A_DARK : 0;
      for (int k = c.lane; k < n; k += 32) {
        int x = rm.x0 + k % w, y = rm.y0 + k / w;
        bool he = (y == rm.y0 || y == rm.y1 - 1), ve = (x == rm.x0 || x == rm.x1 - 1);
        S[y * W + x] = he ? S_WALLX : (ve ? S_WALLY : S_FLOOR);
        A[y * W + x] = (he || ve) ? 0 : fl;
      }
      __syncwarp();
    } else if (rm.kind == K_MAZE) {  // UNIFORM: each passage cell rolls gen_attr in index order
      // the marked cells are found 32 at a time (ballot); the draws stay serial, in ascending cell order
      const uint32_t dark_level = P.dark_level, hidden_inv = P.hidden_passage_rate_inv;  // not re-read per cell
      const int w = rm.x1 - rm.x0, total = w * (rm.y1 - rm.y0);
      for (int base = 0; base < total; base += 32) {
        const int k = base + c.lane;
        int idx = -1;
        if (k < total) idx = (rm.y0 + k / w) * W + rm.x0 + k % w;
        uint32_t bal = __ballot_sync(RG_FULL, idx >= 0 && (A[idx] & A_MARK));
        while (bal) {
          const int cell = __shfl_sync(RG_FULL, idx, __ffs(bal) - 1);
          bal &= bal - 1;
          uint8_t attr = 0;
          if (r.range32G(0, dark_level) < level && r.does_happenG(hidden_inv)) attr = A_HIDDEN;
          S[cell] = S_PASSAGE;
          A[cell] = A_MARK | attr;
        }
        __syncwarp();
      }
    }
  }
}

// floor.rs:85-102: one registered passage cell; `ra` is the attribute stream of the replay
struct PassageRates {  // read once per connection, not once per cell (the sampler calls in between are opaque)
  uint32_t dark_level, locked_inv, hidden_inv;
};
RG_DEV void apply_passage_cell(Ctx& c, Rng& ra, int x, int y, uint8_t surface, uint32_t level, const PassageRates& pr) {
  RG_PLANES(c);
  int idx = y * c.W + x;
  if (!inb(c, x, y)) {
    set_panic(c);
    return;
  }
  uint8_t keep = A[idx] & (A_DOOR | A_MARK);
  uint8_t attr = 0;
  if (surface == S_DOOR) {
    keep |= A_DOOR;
    if (ra.range32G(0, pr.dark_level) < level && ra.does_happenG(pr.locked_inv)) attr = A_LOCKED;
  } else {
    if (ra.range32G(0, pr.dark_level) < level && ra.does_happenG(pr.hidden_inv)) attr = A_HIDDEN;
  }
  A[idx] = keep | attr;
  if (!attr) S[idx] = surface;
}

// passages::select_start_or_end + edges passages.rs:143-219
__device__ void select_start_or_end(Ctx& c, Rng& r, int room, int d, int& ox, int& oy) {
  RG_PLANES(c);
  RoomD rm = st->rooms[room];
  if (rm.kind == K_NORMAL) {  // edges(range, d, inclusive) then SliceRandom::choose (usize lane)
    if (d == D_DOWN || d == D_UP) {
      int len = rm.x1 - rm.x0 - 2;
      int i = (int)r.range64G(0, (uint64_t)len);
      ox = rm.x0 + 1 + i;
      oy = (d == D_DOWN) ? rm.y1 - 1 : rm.y0;
    } else {
      int len = rm.y1 - rm.y0 - 2;
      int i = (int)r.range64G(0, (uint64_t)len);
      oy = rm.y0 + 1 + i;
      ox = (d == D_RIGHT) ? rm.x1 - 1 : rm.x0;
    }
    return;
  }
  if (rm.kind == K_EMPTY) {
    ox = rm.x0;
    oy = rm.y0;
    return;
  }
  // maze: shrink from the far side until an edge line holds a passage cell (passages.rs:149-176)
  int x0 = rm.x0, y0 = rm.y0, x1 = rm.x1, y1 = rm.y1;
  const int W = c.W;
  while (x0 < x1 && y0 < y1) {
    bool horiz = (d == D_DOWN || d == D_UP);
    int fix = (d == D_DOWN) ? y1 - 1 : (d == D_UP) ? y0 : (d == D_RIGHT) ? x1 - 1 : x0;
    int lo = horiz ? x0 : y0, hi = horiz ? x1 : y1;
    // PARALLEL: the passage cells on the line, 32 positions per ballot (a line is at most 160 cells long)
    uint32_t bal[5] = {0, 0, 0, 0, 0};
    int cnt = 0;
    for (int q = 0; q < 5 && lo + 32 * q < hi; ++q) {
      const int t = lo + 32 * q + c.lane;
      const int x = horiz ? t : fix, y = horiz ? fix : t;
      // Maze::has_cd tests membership in the ORIGINAL range (maze.rs:25-31)
      const bool member = t < hi && in_rect(rm, x, y) && (A[y * W + x] & A_MARK);
      bal[q] = __ballot_sync(RG_FULL, member);
      cnt += __popc(bal[q]);
    }
    if (cnt) {
      int n = (int)r.range64G(0, (uint64_t)cnt);
      for (int q = 0; q < 5; ++q) {
        const int pc = __popc(bal[q]);
        if (n < pc) {
          const int t = lo + 32 * q + nth_set_bit(bal[q], (uint32_t)n);
          ox = horiz ? t : fix;
          oy = horiz ? fix : t;
          return;
        }
        n -= pc;
      }
    }
    if (d == D_DOWN) --y1;
    else if (d == D_LEFT) --x0;
    else if (d == D_RIGHT) --x1;
    else --y0;
    if (x0 < 0 || y0 < 0) break;
  }
  set_panic(c);  // "cannot find maze floor"
  ox = rm.x0;
  oy = rm.y0;
}

// passages::connect_2rooms passages.rs:84-133, first half: where the passage starts, ends and turns (three
// draws on the dungeon stream). Which cells it covers follows from that alone, so the connection is only
// RECORDED here; the cells are laid - and their attributes rolled on the second stream - afterwards, in the
// same order (apply_connections), exactly as the reference registers cells while digging and rolls
// gen_attr when the digging is done (floor.rs:73-102). A pair of adjacent rooms is connected at most once,
// so there are at most 2*nx*ny - nx - ny <= 24 records.
struct ConnList {
  uint64_t rec[32];
  int n;
};
__device__ void connect_2rooms(Ctx& c, Rng& rd, ConnList& cl, int r1, int r2, int d) {
  RG_PLANES(c);
  if (d == D_UP || d == D_LEFT) {
    int t = r1; r1 = r2; r2 = t;
    d = reverse_dir(d);
  }
  int sx, sy, ex, ey;
  select_start_or_end(c, rd, r1, d, sx, sy);
  select_start_or_end(c, rd, r2, reverse_dir(d), ex, ey);
  int turn;
  if (d == D_DOWN) {
    if (!(sy + 1 < ey)) { set_panic(c); return; }
    turn = rd.range_i32G(sy + 1, ey);
  } else {
    if (!(sx + 1 < ex)) { set_panic(c); return; }
    turn = rd.range_i32G(sx + 1, ex);
  }
  if (cl.n >= 32) { set_panic(c); return; }
  const uint64_t doors = (st->rooms[r1].kind == K_NORMAL ? 1ull : 0ull) | (st->rooms[r2].kind == K_NORMAL ? 2ull : 0ull);
  cl.rec[cl.n++] = (uint64_t)sx | ((uint64_t)sy << 8) | ((uint64_t)ex << 16) | ((uint64_t)ey << 24) |
                   ((uint64_t)turn << 32) | ((uint64_t)(d == D_DOWN ? 1 : 0) << 40) | (doors << 41);
}
// passages.rs:98-133, second half, for every recorded connection in order: the two ends, then the three legs.
__device__ void apply_connections(Ctx& c, Rng& ra, const ConnList& cl, uint32_t level) {
  const PassageRates pr = {c.P->dark_level, c.P->locked_door_rate_inv, c.P->hidden_passage_rate_inv};
  for (int i = 0; i < cl.n && !c.panic; ++i) {
    const uint64_t r = cl.rec[i];
    const int sx = (int)(r & 0xFF), sy = (int)((r >> 8) & 0xFF), ex = (int)((r >> 16) & 0xFF), ey = (int)((r >> 24) & 0xFF);
    const int turn = (int)((r >> 32) & 0xFF);
    const int d = ((r >> 40) & 1) ? D_DOWN : D_RIGHT;
    apply_passage_cell(c, ra, sx, sy, ((r >> 41) & 1) ? S_DOOR : S_PASSAGE, level, pr);
    apply_passage_cell(c, ra, ex, ey, ((r >> 42) & 1) ? S_DOOR : S_PASSAGE, level, pr);
    int tsx, tsy, tex, tey, tdir;
    if (d == D_DOWN) {
      tdir = (sx < ex) ? D_RIGHT : D_LEFT;
      tsx = sx; tsy = turn; tex = ex; tey = turn;
    } else {
      tdir = (sy < ey) ? D_DOWN : D_UP;
      tsx = turn; tsy = sy; tex = turn; tey = ey;
    }
    int guard = c.W + c.H + 4;
    int x = sx + ddx(d), y = sy + ddy(d);  // .skip(1)
    for (; (x != tsx || y != tsy) && guard > 0; x += ddx(d), y += ddy(d), --guard)
      apply_passage_cell(c, ra, x, y, S_PASSAGE, level, pr);
    for (x = tsx, y = tsy; (x != tex || y != tey) && guard > 0; x += ddx(tdir), y += ddy(tdir), --guard)
      apply_passage_cell(c, ra, x, y, S_PASSAGE, level, pr);
    for (x = tex, y = tey; (x != ex || y != ey) && guard > 0; x += ddx(d), y += ddy(d), --guard)
      apply_passage_cell(c, ra, x, y, S_PASSAGE, level, pr);
    if (guard <= 0) set_panic(c);
  }
}

// Node::candidates (passages.rs:252-262) of room `a` in ascending room id - the order in which
// select_candidate (passages.rs:69-82) meets them: Up (a-nx) < Left (a-1) < Right (a+1) < Down (a+nx).
// Fixed slots (0 Up, 1 Left, 2 Right, 3 Down) with a validity flag instead of a packed list: every index is a
// compile-time constant after unrolling, so the struct lives in registers (a packed list is indexed
// dynamically while it is built and ends up in local memory).
struct Neigh {
  int id[4];
  bool ok[4];
};
RG_DEV int neigh_dir(int slot) { return slot == 0 ? D_UP : slot == 1 ? D_LEFT : slot == 2 ? D_RIGHT : D_DOWN; }
RG_DEV Neigh neighbours(const Ctx& c, int a) {
  Neigh r;
  const int ay = a / c.nx, ax = a - ay * c.nx;
  r.ok[0] = ay > 0;          r.id[0] = a - c.nx;
  r.ok[1] = ax > 0;          r.id[1] = a - 1;
  r.ok[2] = ax < c.nx - 1;   r.id[2] = a + 1;
  r.ok[3] = ay < c.ny - 1;   r.id[3] = a + c.nx;
  return r;
}

// passages::dig_passges passages.rs:16-67: which rooms get connected, in which order (dungeon stream only).
__device__ void dig_passages(Ctx& c, Rng& rd, ConnList& cl) {
  const int n = c.nrooms;
  uint64_t conn = 0;  // bit room*4+dir : connected to the neighbour in that direction
  uint32_t selected;
  int cur = (int)rd.range64G(0, (uint64_t)n);
  selected = 1u << cur;
  int guard = 4096;
  while (__popc(selected) < n && --guard > 0) {
    int pick = -1, pdir = 0;
    uint32_t k = 0;
    const Neigh nb = neighbours(c, cur);
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // select_candidate passages.rs:69-82
      if (!nb.ok[q] || ((selected >> nb.id[q]) & 1u)) continue;
      if (rd.does_happenG(k + 1)) { pick = nb.id[q]; pdir = neigh_dir(q); }
      ++k;
    }
    if (pick >= 0) {
      selected |= 1u << pick;
      conn |= 1ull << (cur * 4 + pdir);
      conn |= 1ull << (pick * 4 + reverse_dir(pdir));
      connect_2rooms(c, rd, cl, cur, pick, pdir);
    } else {
      uint32_t cnt = __popc(selected);
      cur = nth_set_bit(selected, (uint32_t)rd.range64G(0, cnt));
    }
    if (c.panic) return;
  }
  if (guard <= 0) { set_panic(c); return; }
  uint32_t try_num = rd.range32G(0, c.P->max_extra_edges);
  for (uint32_t t = 0; t < try_num; ++t) {
    int room1 = (int)rd.range64G(0, (uint64_t)n);
    int pick = -1, pdir = 0;
    uint32_t k = 0;
    const Neigh nb = neighbours(c, room1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!nb.ok[q] || ((conn >> (room1 * 4 + neigh_dir(q))) & 1ull)) continue;
      if (rd.does_happenG(k + 1)) { pick = nb.id[q]; pdir = neigh_dir(q); }
      ++k;
    }
    if (pick >= 0) {
      conn |= 1ull << (room1 * 4 + pdir);
      conn |= 1ull << (pick * 4 + reverse_dir(pdir));
      connect_2rooms(c, rd, cl, room1, pick, pdir);
    }
    if (c.panic) return;
  }
}

// Floor::gen_floor floor.rs:50-104 up to and including the passages: the part of a level that depends on the
// dungeon stream, the level number and the config only (no items, monsters, player). Leaves A_MARK on maze cells.
__device__ void gen_skeleton(Ctx& c, Rng& rd, uint32_t level) {
  RG_PLANES(c);
  __syncwarp();
  for (int ch = c.lane; ch < c.CP / 16; ch += 32) {  // fresh field
    *reinterpret_cast<uint4*>(S + ch * 16) = make_uint4(0x07070707u, 0x07070707u, 0x07070707u, 0x07070707u);
    *reinterpret_cast<uint4*>(A + ch * 16) = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  gen_rooms(c, rd, level);
  lay_rooms(c, rd, level);
  // dig (dungeon stream), then lay the recorded passages while a second stream, which starts where the
  // dig ended, rolls the cell attributes (connect_2rooms / apply_connections)
  ConnList cl;
  cl.n = 0;
  dig_passages(c, rd, cl);
  Rng ra = rd;
  apply_connections(c, ra, cl, level);
  rd = ra;
}

// Room::select_cell rooms.rs:132-144 over the implicit set "free cells of this room":
// interior floor cells (normal) or marked cells (maze), minus at most one occupied cell
// (`excl`, a cell index or -1). During generation a room's set never loses more than one
// member before it is sampled (see DESIGN.md "implicit cell sets").
__device__ int select_in_room(Ctx& c, Rng& r, int room, int excl) {
  RG_PLANES(c);
  RoomD rm = st->rooms[room];
  const int W = c.W;
  if (rm.kind == K_EMPTY) return -1;
  if (rm.kind == K_NORMAL) {
    int iw = rm.x1 - rm.x0 - 2, ih = rm.y1 - rm.y0 - 2;
    int cnt = iw * ih - (excl >= 0 ? 1 : 0);
    if (cnt <= 0) return -1;
    int n = (int)r.range64G(0, (uint64_t)cnt);  // FenwickSet::select (usize lane) fenwick.rs:90-96
    if (excl >= 0) {
      int eo = (excl % W - rm.x0 - 1) + (excl / W - rm.y0 - 1) * iw;
      if (n >= eo) ++n;
    }
    return (rm.y0 + 1 + n / iw) * W + rm.x0 + 1 + n % iw;
  }
  int cnt = (int)rm.ncells - (excl >= 0 ? 1 : 0);
  if (cnt <= 0) return -1;
  int n = (int)r.range64G(0, (uint64_t)cnt);
  // PARALLEL: the n-th marked cell in row-major order, 32 cells of the maze's rect per trip (one ballot,
  // one popcount) instead of one cell per trip
  const int w = rm.x1 - rm.x0, total = w * (rm.y1 - rm.y0);
  for (int base = 0; base < total; base += 32) {
    const int k = base + c.lane;
    int idx = -1;
    bool member = false;
    if (k < total) {
      idx = (rm.y0 + k / w) * W + rm.x0 + k % w;
      member = (A[idx] & A_MARK) && idx != excl;
    }
    const uint32_t bal = __ballot_sync(RG_FULL, member);
    const int pc = __popc(bal);
    if (n < pc) return __shfl_sync(RG_FULL, idx, nth_set_bit(bal, (uint32_t)n));
    n -= pc;
  }
  set_panic(c);
  return -1;
}

// Floor::select_cell floor.rs:333-346. excl_kind: 0 = occupied by this room's gold (object set),
// 1 = occupied by this room's monster (character set)
__device__ int select_in_floor(Ctx& c, Rng& r, int excl_kind) {
  RG_PLANES(c);
  uint32_t cand = 0;
  for (int i = 0; i < c.nrooms; ++i)
    if (st->rooms[i].kind != K_EMPTY) cand |= 1u << i;
  while (cand) {
    uint32_t cnt = __popc(cand);
    int room = nth_set_bit(cand, (uint32_t)r.range64G(0, cnt));
    int excl = -1;
    if (excl_kind == 0) {
      if (st->item_pos[room] != 0xFFFF) excl = st->item_pos[room];
    } else {
      MonD m = st->mon[room];
      if (m.flags & MF_PRESENT) excl = m.y * c.W + m.x;
    }
    int pos = select_in_room(c, r, room, excl);
    if (pos >= 0) return pos;
    cand &= ~(1u << room);
  }
  return -1;
}

// EnemyHandler::gen_enemy / select / exp_add enemies.rs:265-320
__device__ bool gen_enemy(Ctx& c, Rng& re, uint32_t rmin, uint32_t rmax, bool has_gold, MonD& out) {
  const rg_params& P = *c.P;
  if (!re.parcentG(has_gold ? P.appear_rate_gold : P.appear_rate_nogold)) return false;
  uint32_t len = P.n_enemies;
  uint32_t idx = re.range32G(rmin, rmax);
  if (idx > len) {
    uint32_t rr = len < 5 ? len : 5;
    idx = (uint32_t)re.range64G(len - rr, len);
  }
  if (idx >= len) return false;
  const rg_enemy_kind& k = P.enemies[idx];
  int la = (int)lev_add(c);
  int lvl = k.level + la;
  int hp = 0;
  for (int i = 0; i < 8; ++i) hp += (int)re.range64G(1, (uint64_t)(lvl + 1));
  int base = (lvl == 1) ? hp / 8 : hp / 6;
  uint32_t add = (10 <= lvl) ? (uint32_t)base * 20u : (uint32_t)base * 4u;
  out.kind = (uint8_t)idx;
  out.flags = MF_PRESENT;
  out.hp = hp;
  out.exp = k.exp + (uint32_t)(la * 10) + add;
  return true;
}

#else
// rogue::Dungeon::new_level_ rogue/mod.rs:434-481 + Floor::gen_floor floor.rs:50-104
// + setup_items :132-153 + setup_stair :156-167 + place_enemies :106-130,
// then actions::new_level's player placement (actions.rs:134-137).
RG_GEN_NEW_LEVEL_ATTR void new_level(Ctx* cp, bool is_initial) {
  Ctx& c = *cp;
  RG_PLANES(c);
  const rg_params& P = *c.P;
  const int W = c.W;
  if (!is_initial) {
    snapshot_suspended_maps(c);  // suspended DistCache maps belong to the floor that is about to be replaced
    // The descending step shows the visited map of the floor being left:
    // emit it now, before the planes are overwritten.
    __syncwarp();
    for (int ch = c.lane; ch < c.CP / 16; ch += 32) {
      uint4 a = *reinterpret_cast<const uint4*>(A + ch * 16);
      uint32_t v[4] = {a.x, a.y, a.z, a.w};
      uint32_t bits = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) bits |= ((v[k >> 2] >> (8 * (k & 3))) & 1u) << k;
      reinterpret_cast<uint16_t*>(c.g_hist)[ch] = (uint16_t)bits;
    }
    c.hist_done = 1;
  }
  st->level += 1;
  st->dirty_rows = ~0ull;  // a new floor: everything is recomposed
  st->spec_req = 0;
  const uint32_t level = (uint32_t)st->level;
  Rng rd = c.rd;
  // rooms, mazes, passages and their attributes: a function of (dungeon stream, level, config) alone. A descent
  // first looks for that part built ahead of time by k_spec_build from the same stream state (take_spec).
  if (is_initial || !take_spec(c, rd, level)) gen_skeleton(c, rd, level);
  // Floor::setup_items: cell from the dungeon stream, amount from the item stream (gold.rs:18-24)
  Rng ri = c.ri;
  for (int i = 0; i < MAX_ROOMS; ++i) {  // every slot: k_step_fast looks at all of them
    st->item_pos[i] = 0xFFFF;
    st->item_amt[i] = 0;
  }
  for (int i = 0; i < c.nrooms; ++i) {
    int pos = select_in_room(c, rd, i, -1);
    if (pos < 0) continue;
    if (!ri.does_happenG(P.gold_rate_inv)) continue;
    uint32_t num = ri.range32G(0, P.gold_base + P.gold_per_level * level) + P.gold_minimum;
    st->item_pos[i] = (uint16_t)pos;
    st->item_amt[i] = num;
    st->rooms[i].flags |= RF_GOLD;
  }
  c.ri = ri;
  {  // Floor::setup_stair
    int pos = select_in_floor(c, rd, 0);
    if (pos < 0) set_panic(c);
    else S[pos] = S_STAIR;
    st->stair_pos = (uint16_t)(pos < 0 ? 0xFFFF : pos);
  }
  for (int i = 0; i < MAX_ROOMS; ++i) st->mon[i].flags = 0;  // remove_enemies (a fresh handler is empty too)
  if (P.n_enemies != 0) {  // Floor::place_enemies
    Rng re = c.re;
    uint32_t mn = level >= 4 ? level - 4 : 0, mx = level + 6;
    for (int i = 0; i < c.nrooms; ++i) {
      int pos = select_in_room(c, rd, i, -1);
      if (pos < 0) continue;
      MonD m;
      if (gen_enemy(c, re, mn, mx, (st->rooms[i].flags & RF_GOLD) != 0, m)) {
        m.x = (uint8_t)(pos % W);
        m.y = (uint8_t)(pos / W);
        st->mon[i] = m;
      }
    }
    c.re = re;
  }
  if (!P.hide_dungeon) {  // rogue/mod.rs:465-475
    __syncwarp();
    for (int k = W + c.lane; k < (c.H - 1) * W; k += 32) A[k] |= A_VISIBLE;
    __syncwarp();
  }
  if (is_initial) {  // Player::init_items -> weapon.rs:159 draws on the item stream (core/src/lib.rs:206-207)
    Rng r2 = c.ri;
    for (uint32_t i = 0; i < P.n_init_draws; ++i) r2.range32G(P.init_draw_lo[i], P.init_draw_hi[i]);
    c.ri = r2;
  }
  int ppos = select_in_floor(c, rd, 1);
  c.rd = rd;
  if (ppos < 0) { set_panic(c); ppos = W + 1; }
  st->px = (int16_t)(ppos % W);
  st->py = (int16_t)(ppos / W);
  __syncwarp();
  for (int k = c.lane; k < c.CP / 16; k += 32) {
    uint4 v = reinterpret_cast<uint4*>(A)[k];
    v.x &= 0x7F7F7F7Fu; v.y &= 0x7F7F7F7Fu; v.z &= 0x7F7F7F7Fu; v.w &= 0x7F7F7F7Fu;
    reinterpret_cast<uint4*>(A)[k] = v;
  }
  __syncwarp();
  build_walk(c);
  player_in(c, st->px, st->py, true);
  c.a_dirty = 1;
  c.s_dirty = 1;
}

#endif
