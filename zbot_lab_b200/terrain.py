"""Height-field terrain of the rough manager-based task (``zbot-6b-walking-m-rough-v0``).

The reference's ``ZbotLabRoughEnvCfg`` imports its ground from Isaac Lab: ``TerrainImporterCfg(terrain_type="generator",
terrain_generator=ROUGH_TERRAINS_CFG, max_init_terrain_level=5)`` (``/root/reference/source/zbot/zbot/tasks/zbotlab_manager/
zbotlab_env_cfg.py:31, 44-48``) and walks the robots over it with the ``terrain_levels_vel`` curriculum
(``mdp/curriculums.py:26-55``).  ``ROUGH_TERRAINS_CFG``, the sub-terrain generators and ``TerrainImporter`` are
**[IL-upstream]**: not vendored by the reference, version unpinned, and the generators draw from numpy's global stream --
so this module is a RESTATEMENT FROM UPSTREAM KNOWLEDGE (same tile grid, same sub-terrain families, proportions, parameter
ranges and difficulty schedule), not a bit-level twin; ``DESIGN.md`` §3 says so.  What is exact is what the reference itself
contains: the curriculum rule and the env-origin bookkeeping it drives.

Representation: ONE float32 height field ``H[ix][iy]`` over the whole tile grid plus a flat border (0.1 m cells, the
generator's ``horizontal_scale``), sampled bilinearly by the contact code of the fused kernel
(``csrc/zbot_core.h: TerrainGround``).  Contact normals are vertical: a height field has no overhangs, and the risers of a
stair are met as a steep ramp one cell wide.  Meshes / trimesh are not used.
"""
from __future__ import annotations

import math

import numpy as np

from .utils.configclass import Cfg

F = np.float32


class SubTerrainCfg(Cfg):
    kind = "flat"
    proportion = 1.0
    params = {}

    def __init__(self, kind="flat", proportion=1.0, **params):
        super().__init__(kind=kind, proportion=float(proportion), params=dict(params))


class TerrainGeneratorCfg(Cfg):
    """isaaclab.terrains.TerrainGeneratorCfg [IL-upstream], the fields ROUGH_TERRAINS_CFG sets."""
    size = (8.0, 8.0)
    border_width = 20.0
    num_rows = 10              # terrain LEVELS (difficulty grows with the row when curriculum = True)
    num_cols = 20              # terrain TYPES
    horizontal_scale = 0.1
    vertical_scale = 0.005
    slope_threshold = 0.75
    curriculum = False
    seed = None
    sub_terrains = {}


def rough_terrains_cfg() -> TerrainGeneratorCfg:
    """ROUGH_TERRAINS_CFG of isaaclab.terrains.config.rough [IL-upstream], restated."""
    c = TerrainGeneratorCfg()
    c.sub_terrains = {
        "pyramid_stairs": SubTerrainCfg("pyramid_stairs", 0.2, step_height_range=(0.05, 0.23), step_width=0.3, platform_width=3.0, inverted=False),
        "pyramid_stairs_inv": SubTerrainCfg("pyramid_stairs", 0.2, step_height_range=(0.05, 0.23), step_width=0.3, platform_width=3.0, inverted=True),
        "boxes": SubTerrainCfg("random_grid", 0.2, grid_width=0.45, grid_height_range=(0.05, 0.2), platform_width=2.0),
        "random_rough": SubTerrainCfg("random_uniform", 0.2, noise_range=(0.02, 0.10), noise_step=0.02),
        "hf_pyramid_slope": SubTerrainCfg("pyramid_slope", 0.1, slope_range=(0.0, 0.4), platform_width=2.0, inverted=False),
        "hf_pyramid_slope_inv": SubTerrainCfg("pyramid_slope", 0.1, slope_range=(0.0, 0.4), platform_width=2.0, inverted=True),
    }
    return c


# ------------------------------------------------------------------------------------------------ sub-terrain height fields
def _ring_index(n: int) -> np.ndarray:
    """distance (in cells) of every cell of an n x n tile from the tile border"""
    i = np.arange(n)
    d = np.minimum(i, n - 1 - i)
    return np.minimum(d[:, None], d[None, :])


def _lerp(rng_, difficulty):
    return rng_[0] + difficulty * (rng_[1] - rng_[0])


def sub_terrain(kind: str, p: dict, difficulty: float, n: int, cell: float, rng: np.random.Generator) -> np.ndarray:
    """(n, n) float64 heights of one tile; the tile CENTRE is where a robot is spawned."""
    ring = _ring_index(n).astype(np.float64)
    if kind == "flat":
        return np.zeros((n, n))
    if kind == "pyramid_stairs":
        h = _lerp(p["step_height_range"], difficulty)
        w = max(1, int(round(p["step_width"] / cell)))
        plat = int(round(p["platform_width"] / cell))
        steps_max = max(0, (n - plat) // 2 // w)
        z = np.minimum(np.floor(ring / w), steps_max) * h
        return -z if p.get("inverted") else z
    if kind == "pyramid_slope":
        s = _lerp(p["slope_range"], difficulty)
        plat = int(round(p["platform_width"] / cell))
        top = max(0, (n - plat) // 2)
        z = np.minimum(ring, top) * cell * s
        return -z if p.get("inverted") else z
    if kind == "random_grid":
        hgt = _lerp(p["grid_height_range"], difficulty)
        g = max(1, int(round(p["grid_width"] / cell)))
        m = -(-n // g)
        boxes = rng.uniform(-hgt, hgt, (m, m))
        z = np.kron(boxes, np.ones((g, g)))[:n, :n]
        plat = int(round(p["platform_width"] / cell))
        a = (n - plat) // 2
        z[a:n - a, a:n - a] = 0.0
        return z
    if kind == "random_uniform":
        lo, hi = p["noise_range"]
        step = p["noise_step"]
        levels = np.arange(lo, hi + 0.5 * step, step)
        coarse = rng.choice(levels, (n // 2 + 2, n // 2 + 2)) * (0.25 + 0.75 * difficulty)     # sampled every 2 cells, interpolated
        x = np.arange(n) / 2.0
        i0 = np.floor(x).astype(int)
        t = x - i0
        rows = coarse[i0] * (1 - t)[:, None] + coarse[i0 + 1] * t[:, None]
        z = rows[:, i0] * (1 - t)[None, :] + rows[:, i0 + 1] * t[None, :]
        return z - z[n // 2, n // 2]
    raise NotImplementedError(f"sub-terrain kind {kind!r}")


class Terrain:
    """The generated ground: height field + tile origins + the env-origin bookkeeping of ``TerrainImporter``."""

    def __init__(self, cfg: TerrainGeneratorCfg, seed: int = 0):
        self.cfg = cfg
        rng = np.random.default_rng(seed if cfg.seed is None else cfg.seed)
        cell = float(cfg.horizontal_scale)
        n = int(round(cfg.size[0] / cell))
        assert abs(cfg.size[0] - cfg.size[1]) < 1e-9, "square tiles"
        rows, cols = int(cfg.num_rows), int(cfg.num_cols)
        b = int(round(cfg.border_width / cell))
        self.cell, self.tile_cells, self.rows, self.cols, self.border_cells = cell, n, rows, cols, b
        nx, ny = rows * n + 2 * b + 1, cols * n + 2 * b + 1
        H = np.zeros((nx, ny))
        # columns -> sub-terrain kinds by cumulative proportion (TerrainGenerator._generate_curriculum_terrains [IL-upstream])
        names = list(cfg.sub_terrains)
        prop = np.array([cfg.sub_terrains[k].proportion for k in names], float)
        cum = np.cumsum(prop / prop.sum())
        self.col_kind = [names[min(int(np.searchsorted(cum, (c + 0.5) / cols)), len(names) - 1)] for c in range(cols)]
        self.origins = np.zeros((rows, cols, 3), F)
        # world frame: the grid is centred on (0, 0); tile (r, c) spans x in [x0 + r L, x0 + (r+1) L)
        L = cfg.size[0]
        self.x0, self.y0 = -0.5 * rows * L - b * cell, -0.5 * cols * L - b * cell      # world coordinate of H[0][0]
        for r in range(rows):
            for c in range(cols):
                st = cfg.sub_terrains[self.col_kind[c]]
                lo, hi = (r / rows, (r + 1) / rows) if cfg.curriculum else (0.0, 1.0)
                diff = float(rng.uniform(lo, hi))
                z = sub_terrain(st.kind, st.params, diff, n, cell, rng)
                z = np.round(z / cfg.vertical_scale) * cfg.vertical_scale
                H[b + r * n:b + (r + 1) * n, b + c * n:b + (c + 1) * n] = z
                ci = b + r * n + n // 2, b + c * n + n // 2
                self.origins[r, c] = (self.x0 + ci[0] * cell, self.y0 + ci[1] * cell, H[ci])
        self.heights = np.ascontiguousarray(H, F)

    # --------------------------------------------------------------------------------------------- sampling (host twin of the kernel's)
    def height_at(self, x, y, dtype=np.float32):
        """Bilinear sample at world (x, y) -- the arithmetic of csrc/zbot_core.h: TerrainGround in `dtype` (the grid and its
        geometry parameters are float32 either way, as the kernel receives them)."""
        D = dtype
        H = self.heights
        inv_cell = D(F(1.0) / F(self.cell))
        fx = (np.asarray(x, D) - D(F(self.x0))) * inv_cell
        fy = (np.asarray(y, D) - D(F(self.y0))) * inv_cell
        fx = np.clip(fx, D(0), D(H.shape[0]) - D(1.001)).astype(D)
        fy = np.clip(fy, D(0), D(H.shape[1]) - D(1.001)).astype(D)
        ix, iy = fx.astype(np.int64), fy.astype(np.int64)
        tx, ty = (fx - ix).astype(D), (fy - iy).astype(D)
        h00, h10, h01, h11 = (H[ix, iy].astype(D), H[ix + 1, iy].astype(D), H[ix, iy + 1].astype(D), H[ix + 1, iy + 1].astype(D))
        a = (h00 + tx * (h10 - h00)).astype(D)
        b = (h01 + tx * (h11 - h01)).astype(D)
        return (a + ty * (b - a)).astype(D)

    # --------------------------------------------------------------------------------------------- TerrainImporter bookkeeping [IL-upstream]
    def initial_levels_types(self, num_envs: int, max_init_level, rng: np.random.Generator):
        """``TerrainImporter._compute_env_origins_curriculum``: levels ~ randint(0, max_init + 1), types = floor(i / (N / cols))."""
        hi = self.rows - 1 if max_init_level is None else min(int(max_init_level), self.rows - 1)
        levels = rng.integers(0, hi + 1, num_envs)
        types = np.floor(np.arange(num_envs) / (num_envs / self.cols)).astype(np.int64).clip(0, self.cols - 1)
        return levels, types

    def update_levels(self, levels, move_up, move_down, rand_level):
        """``TerrainImporter.update_env_origins``: +1 / -1, a robot that solves the last level restarts at a random one."""
        lv = levels + move_up.astype(np.int64) - move_down.astype(np.int64)
        return np.where(lv >= self.rows, rand_level, np.clip(lv, 0, None))


def terrain_levels_vel(root_xy_local, command_xy, tile_size, episode_s):
    """``mdp/curriculums.py:26-55`` on env-local root positions (= root_pos_w - env_origin): the move_up / move_down masks."""
    dist = np.linalg.norm(np.asarray(root_xy_local, F), axis=1).astype(F)
    up = dist > F(tile_size / 2)
    down = dist < (np.linalg.norm(np.asarray(command_xy, F), axis=1).astype(F) * F(episode_s) * F(0.5))
    down &= ~up
    return up, down
