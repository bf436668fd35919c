"""ctypes mirror of include/sk_engine.h and a thin object wrapper around the C ABI.

The wrapper is parametrised by (shared library, symbol prefix) so that the test-suite can drive the CPU
oracle (prefix ``sko_``) through exactly the same Python code as the CUDA engine (prefix ``sk_engine_``).
The product never loads the oracle: :func:`load_engine_library` only knows ``libskirt9_b200.so`` and fails
loudly when it has not been built.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ENGINE_LIB_PATH = os.path.join(_HERE, "csrc", "libskirt9_b200.so")

SK_OK, SK_ERR_INVALID, SK_ERR_UNSUPPORTED, SK_ERR_CUDA, SK_ERR_STATE = range(5)

SK_SRC_POINT, SK_SRC_GEOMETRIC = 1, 2
SK_GEOM_NONE, SK_GEOM_SHELL, SK_GEOM_EXPDISK, SK_GEOM_RING, SK_GEOM_SPIRAL_EXPDISK = range(5)
SK_SED_TABULATED, SK_SED_BLACKBODY = 1, 2
SK_BIAS_NONE, SK_BIAS_LOGUNIFORM, SK_BIAS_OLIGO = 0, 1, 2
SK_VEL_NONE, SK_VEL_CONSTANT, SK_VEL_RADIAL, SK_VEL_CYLINDRICAL = 0, 1, 2, 3
SK_INSTR_SED, SK_INSTR_FRAME, SK_INSTR_FULL = 1, 2, 3
(SK_COMP_TOTAL, SK_COMP_TRANSPARENT, SK_COMP_PRIMARY_DIRECT, SK_COMP_PRIMARY_SCATTERED, SK_COMP_SECONDARY_DIRECT,
 SK_COMP_SECONDARY_SCATTERED, SK_COMP_SECONDARY_TRANSPARENT, SK_COMP_PRIMARY_SCATTERED_LEVEL) = range(8)
SK_GEOM_MAX_PARAMS = 12

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class SkConfig(C.Structure):
    _fields_ = [("seed", C.c_uint32), ("force_scattering", C.c_int32), ("min_scatt_events", C.c_int32),
                ("path_length_bias", C.c_double), ("min_weight_reduction", C.c_double), ("device", C.c_int32),
                ("explicit_absorption", C.c_int32)]


class SkWavelengthGrid(C.Structure):
    _fields_ = [("num_bins", C.c_int32), ("num_borders", C.c_int32), ("borders", _dp), ("ell", _ip),
                ("lambda_", _dp), ("dlambda", _dp)]


class SkDustMix(C.Structure):
    _fields_ = [("num_lambda", C.c_int32), ("reserved", C.c_int32), ("lambda_border", _dp), ("sigma_abs", _dp),
                ("sigma_sca", _dp), ("asymmpar", _dp), ("mu", C.c_double)]


class SkSource(C.Structure):
    _fields_ = [("kind", C.c_int32), ("geometry", C.c_int32), ("luminosity", C.c_double),
                ("source_weight", C.c_double), ("position", C.c_double * 3),
                ("geom_params", C.c_double * SK_GEOM_MAX_PARAMS), ("geom_table_n", C.c_int32),
                ("sed_kind", C.c_int32), ("geom_table_x", _dp), ("geom_table_P", _dp), ("sed_n", C.c_int32),
                ("bias_kind", C.c_int32), ("sed_lambda", _dp), ("sed_p", _dp), ("sed_P", _dp),
                ("sed_temperature", C.c_double), ("sed_norm", C.c_double), ("wavelength_bias", C.c_double),
                ("bias_min", C.c_double), ("bias_max", C.c_double), ("oligo_n", C.c_int32), ("reserved", C.c_int32),
                ("oligo_lambda", _dp), ("oligo_probability", C.c_double), ("velocity_kind", C.c_int32),
                ("reserved2", C.c_int32), ("velocity", C.c_double * 3)]


class SkInstrument(C.Structure):
    _fields_ = [("kind", C.c_int32), ("wavelength_grid", C.c_int32), ("inclination", C.c_double),
                ("azimuth", C.c_double), ("roll", C.c_double), ("distance", C.c_double), ("radius", C.c_double),
                ("num_pixels_x", C.c_int32), ("num_pixels_y", C.c_int32), ("field_of_view_x", C.c_double),
                ("field_of_view_y", C.c_double), ("center_x", C.c_double), ("center_y", C.c_double),
                ("record_components", C.c_int32), ("num_scattering_levels", C.c_int32),
                ("record_statistics", C.c_int32), ("reserved", C.c_int32), ("redshift", C.c_double)]


class SkSecondary(C.Structure):
    _fields_ = [("emission_grid", C.c_int32), ("num_temperatures", C.c_int32), ("spatial_bias", C.c_double),
                ("wavelength_bias", C.c_double), ("bias_min", C.c_double), ("bias_max", C.c_double),
                ("temperature", _dp), ("planck_abs", _dp), ("rf_sigma_abs", _dp), ("em_sigma_abs", _dp), ("rf_cmb", _dp)]


SK_DENSITY_MAX_PARAMS = 16


class SkDensityGeometry(C.Structure):
    _fields_ = [("geometry", C.c_int32), ("reserved", C.c_int32), ("number", C.c_double), ("mass", C.c_double),
                ("p", C.c_double * SK_DENSITY_MAX_PARAMS)]

    @classmethod
    def make(cls, geometry, params, number=1.0, mass=1.0):
        g = cls()
        g.geometry, g.number, g.mass = int(geometry), float(number), float(mass)
        if len(params) > SK_DENSITY_MAX_PARAMS:
            raise ValueError("too many density parameters")
        for k, v in enumerate(params):
            g.p[k] = float(v)
        return g


class SkTreePolicy(C.Structure):
    _fields_ = [("min_level", C.c_int32), ("max_level", C.c_int32), ("num_samples", C.c_int32), ("reserved", C.c_int32),
                ("max_dust_fraction", C.c_double), ("max_dust_optical_depth", C.c_double),
                ("max_dust_density_dispersion", C.c_double), ("dust_kappa", C.c_double)]


class SkCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("packets", "forward_paths", "forward_segments", "replay_segments", "peel_paths", "peel_segments",
                 "scatterings", "rf_deposits", "detections", "fallbacks", "kernel_launches", "rounds", "pixel_overflows")] \
        + [("reserved", C.c_uint64 * 3)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_ if n != "reserved"}


class SkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[sk status {code}] {msg}")
        self.code = code


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


def load_engine_library(path: Optional[str] = None) -> C.CDLL:
    """Loads the CUDA engine; there is no fallback of any kind."""
    path = path or os.environ.get("SK_ENGINE_LIB") or ENGINE_LIB_PATH  # SK_ENGINE_LIB: kernel-variant experiments
    if not os.path.exists(path):
        raise RuntimeError(f"{path} has not been built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the engine has no CPU fallback)")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


# names of the C ABI entry points after the prefix; include/sk_engine.h is the authority
ABI_FUNCTIONS = ["create", "destroy", "set_grid_cartesian", "set_grid_octree", "set_grid_voronoi", "set_voronoi_extents",
                 "set_medium", "set_dustmix",
                 "set_wavelength_grids", "set_sources", "set_instruments", "set_secondary", "clear_instruments", "clear_rf",
                 "prepare_primary", "prepare_secondary", "set_history_interleave", "run_segment", "communicate_rf",
                 "absorbed_luminosity",
                 "read_rf", "read_sed", "read_ifu", "read_sed_stats", "read_ifu_stats", "counters"]
SETUP_FUNCTIONS = ["build_octree", "read_octree", "sample_medium", "read_medium"]
ENGINE_ONLY_FUNCTIONS = ["launch_segment", "synchronize", "last_kernel_ms", "last_stage_ms", "device_buffer", "cuda_stream",
                         "measure_gather_peak"]


class Engine:
    """Object wrapper over the C ABI (``sk_engine_*`` by default)."""

    def __init__(self, config: SkConfig, lib: Optional[C.CDLL] = None, prefix: str = "sk_engine_",
                 misc_prefix: str = "sk_"):
        self.lib = lib if lib is not None else load_engine_library()
        self.prefix = prefix
        self._last_error = getattr(self.lib, misc_prefix + "last_error")
        self._last_error.restype = C.c_char_p
        self._h = C.c_void_p()
        self._keep = []
        self.config = config
        self._call("create", C.byref(config), C.byref(self._h))
        self.num_cells = 0
        self.num_nodes = 0
        self.num_rf = 0
        self._instr = []
        self._wlg = []

    # -- plumbing -------------------------------------------------------------------------------
    def _fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f

    def _call(self, name, *args):
        rc = self._fn(name)(*args)
        if rc != SK_OK:
            raise SkError(rc, (self._last_error() or b"").decode())

    def close(self):
        if self._h:
            f = getattr(self.lib, self.prefix + "destroy")
            f.restype = None
            f(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setters --------------------------------------------------------------------------------
    def set_grid_cartesian(self, xv, yv, zv):
        xv, px = _d(xv)
        yv, py = _d(yv)
        zv, pz = _d(zv)
        self._call("set_grid_cartesian", self._h, C.c_int32(len(xv) - 1), C.c_int32(len(yv) - 1),
                   C.c_int32(len(zv) - 1), px, py, pz)

    def set_grid_octree(self, extent, first_child):
        ext, pe = _d(extent)
        fc, pf = _i(first_child)
        self._call("set_grid_octree", self._h, pe, C.c_int32(len(fc)), pf)
        self.num_nodes = len(fc)

    def build_octree(self, extent, policy: SkTreePolicy, media: Sequence[SkDensityGeometry]):
        """DensityTreePolicy::constructTree on the engine's side; returns (num_nodes, num_cells)."""
        ext, pe = _d(extent)
        arr = (SkDensityGeometry * len(media))(*media)
        nn, nc = C.c_uint64(), C.c_uint64()
        self._call("build_octree", self._h, pe, C.byref(policy), C.c_int32(len(media)), arr, C.byref(nn), C.byref(nc))
        self.num_nodes = int(nn.value)
        self.num_grid_cells = int(nc.value)
        return self.num_nodes, self.num_grid_cells

    def read_octree(self):
        out = np.empty(max(self.num_nodes, 1), dtype=np.int32)
        self._call("read_octree", self._h, out.ctypes.data_as(_ip))
        return out[:self.num_nodes]

    def sample_medium(self, medium: SkDensityGeometry, num_samples: int, num_cells: int):
        self._call("sample_medium", self._h, C.byref(medium), C.c_int32(num_samples))
        self.num_cells = num_cells

    def sample_medium_particles(self, particles, density_scale: float, num_samples: int, num_cells: int):
        """ParticleMedium sampled on the engine's side: particles[n][5] = x y z h M (sk_engine_sample_medium_particles)."""
        p, pp = _d(np.asarray(particles, dtype=np.float64).reshape(-1, 5))
        self._call("sample_medium_particles", self._h, C.c_int32(len(p)), pp, C.c_double(density_scale), C.c_int32(num_samples))
        self.num_cells = num_cells

    def read_medium(self):
        dens = np.empty(self.num_cells)
        vol = np.empty(self.num_cells)
        self._call("read_medium", self._h, dens.ctypes.data_as(_dp), vol.ctypes.data_as(_dp))
        return dens, vol

    def set_grid_voronoi(self, extent, sites, nbr_offset, nbr_index):
        ext, pe = _d(extent)
        st, ps = _d(np.asarray(sites, dtype=np.float64).reshape(-1, 3))
        off = np.ascontiguousarray(nbr_offset, dtype=np.int64)
        idx, pi = _i(nbr_index)
        self._call("set_grid_voronoi", self._h, pe, C.c_int32(len(st)), ps, off.ctypes.data_as(C.POINTER(C.c_int64)), pi)

    def build_voronoi(self, extent, sites):
        """VoronoiMeshSnapshot::buildMesh on the engine's side (sk_engine_build_voronoi): returns the number of neighbour entries."""
        ext, pe = _d(extent)
        st, ps = _d(np.asarray(sites, dtype=np.float64).reshape(-1, 3))
        n = C.c_uint64()
        self._call("build_voronoi", self._h, pe, C.c_int32(len(st)), ps, C.byref(n))
        self.num_cells = len(st)
        self._voronoi_entries = int(n.value)
        return int(n.value)

    def read_voronoi(self):
        """(nbr_offset, nbr_index, volumes, boxes) of the tessellation built by build_voronoi."""
        n = self.num_cells
        off = np.empty(n + 1, dtype=np.int64)
        idx = np.empty(self._voronoi_entries, dtype=np.int32)
        vol = np.empty(n)
        box = np.empty((n, 6))
        self._call("read_voronoi", self._h, off.ctypes.data_as(C.POINTER(C.c_int64)), idx.ctypes.data_as(C.POINTER(C.c_int32)),
                   vol.ctypes.data_as(_dp), box.ctypes.data_as(_dp))
        return off, idx, vol, box

    def set_voronoi_extents(self, boxes):
        b, pb = _d(np.asarray(boxes, dtype=np.float64).reshape(-1, 6))
        self._call("set_voronoi_extents", self._h, C.c_int32(len(b)), pb)

    def set_medium(self, number_density, volume=None):
        n, pn = _d(number_density)
        if volume is not None:
            v, pv = _d(volume)
        else:
            pv = None
        self._call("set_medium", self._h, C.c_int32(len(n)), pn, pv)
        self.num_cells = len(n)

    def set_velocities(self, velocity):
        """MediumState::bulkVelocity(m) per cell, [num_cells][3] in m/s; None returns to media at rest."""
        if velocity is None:
            self._call("set_velocities", self._h, C.c_int32(0), None)
            return
        v, pv = _d(np.ascontiguousarray(velocity, dtype=float).reshape(-1))
        self._call("set_velocities", self._h, C.c_int32(len(v) // 3), pv)

    def set_dustmix(self, lambda_border, sigma_abs, sigma_sca, asymmpar, mu):
        lb, p0 = _d(lambda_border)
        sa, p1 = _d(sigma_abs)
        ss, p2 = _d(sigma_sca)
        g, p3 = _d(asymmpar)
        mix = SkDustMix(len(lb), 0, p0, p1, p2, p3, mu)
        self._call("set_dustmix", self._h, C.byref(mix))

    def set_media(self, number_density, volume=None):
        """Several medium components: number_density[h][m] (sk_engine_set_media)."""
        n, pn = _d(np.asarray(number_density, dtype=np.float64).reshape(len(number_density), -1))
        if volume is not None:
            v, pv = _d(volume)
        else:
            pv = None
        self._call("set_media", self._h, C.c_int32(n.shape[1]), C.c_int32(n.shape[0]), pn, pv)
        self.num_cells = n.shape[1]

    def set_dustmixes(self, mixes: Sequence[tuple]):
        """mixes: (lambda_border, sigma_abs, sigma_sca, asymmpar, mu) per medium component (sk_engine_set_dustmixes)."""
        arr = (SkDustMix * len(mixes))()
        keep = []
        for h, (lambda_border, sigma_abs, sigma_sca, asymmpar, mu) in enumerate(mixes):
            k = [_d(lambda_border), _d(sigma_abs), _d(sigma_sca), _d(asymmpar)]
            keep.append(k)
            arr[h] = SkDustMix(len(k[0][0]), 0, k[0][1], k[1][1], k[2][1], k[3][1], mu)
        self._call("set_dustmixes", self._h, C.c_int32(len(mixes)), arr)

    def set_wavelength_grids(self, grids: Sequence[dict], rf_grid: int = -1):
        """grids: dicts with keys borders, ell, lambda, dlambda (see host.DisjointWavelengthGrid.table())."""
        arr = (SkWavelengthGrid * max(1, len(grids)))()
        keep = []
        for k, g in enumerate(grids):
            b, pb = _d(g["borders"])
            e, pe = _i(g["ell"])
            l, pl = _d(g["lambda"])
            d, pd = _d(g["dlambda"])
            keep += [b, e, l, d]
            arr[k] = SkWavelengthGrid(len(l), len(b), pb, pe, pl, pd)
        self._call("set_wavelength_grids", self._h, C.c_int32(len(grids)), arr, C.c_int32(rf_grid))
        self._wlg = [len(g["lambda"]) for g in grids]
        self.num_rf = self._wlg[rf_grid] if rf_grid >= 0 else 0

    def set_sources(self, sources: Sequence[dict], source_bias: float = 0.5):
        arr = (SkSource * len(sources))()
        keep = []
        for k, s in enumerate(sources):
            q = SkSource()
            q.kind = s["kind"]
            q.geometry = s.get("geometry", SK_GEOM_NONE)
            q.luminosity = s["luminosity"]
            q.source_weight = s.get("source_weight", 1.0)
            q.position = (C.c_double * 3)(*s.get("position", (0., 0., 0.)))
            gp = list(s.get("geom_params", ())) + [0.0] * SK_GEOM_MAX_PARAMS
            q.geom_params = (C.c_double * SK_GEOM_MAX_PARAMS)(*gp[:SK_GEOM_MAX_PARAMS])
            if s.get("geom_table_x") is not None:
                a, pa = _d(s["geom_table_x"])
                b, pb = _d(s["geom_table_P"])
                keep += [a, b]
                q.geom_table_n, q.geom_table_x, q.geom_table_P = len(a), pa, pb
            q.sed_kind = s["sed_kind"]
            a, pa = _d(s["sed_lambda"])
            b, pb = _d(s["sed_p"])
            c, pc = _d(s["sed_P"])
            keep += [a, b, c]
            q.sed_n, q.sed_lambda, q.sed_p, q.sed_P = len(a), pa, pb, pc
            q.sed_temperature = s.get("sed_temperature", 0.0)
            q.sed_norm = s.get("sed_norm", 1.0)
            q.wavelength_bias = s.get("wavelength_bias", 0.0)
            q.bias_kind = s.get("bias_kind", SK_BIAS_NONE)
            q.bias_min = s.get("bias_min", 0.0)
            q.bias_max = s.get("bias_max", 0.0)
            if s.get("oligo_lambda") is not None:
                a, pa = _d(s["oligo_lambda"])
                keep.append(a)
                q.oligo_n, q.oligo_lambda = len(a), pa
                q.oligo_probability = s["oligo_probability"]
            q.velocity_kind = s.get("velocity_kind", SK_VEL_NONE)
            q.velocity = (C.c_double * 3)(*s.get("velocity", (0., 0., 0.)))
            arr[k] = q
        self._call("set_sources", self._h, C.c_int32(len(sources)), arr, C.c_double(source_bias))

    def set_instruments(self, instruments: Sequence[dict], has_medium_emission: bool = False):
        arr = (SkInstrument * max(1, len(instruments)))()
        self._instr = []
        for k, s in enumerate(instruments):
            q = SkInstrument()
            q.kind = s["kind"]
            q.wavelength_grid = s.get("wavelength_grid", 0)
            q.inclination = s.get("inclination", 0.0)
            q.azimuth = s.get("azimuth", 0.0)
            q.roll = s.get("roll", 0.0)
            q.distance = s["distance"]
            q.radius = s.get("radius", 0.0)
            q.num_pixels_x = s.get("num_pixels_x", 0)
            q.num_pixels_y = s.get("num_pixels_y", 0)
            q.field_of_view_x = s.get("field_of_view_x", 0.0)
            q.field_of_view_y = s.get("field_of_view_y", 0.0)
            q.center_x = s.get("center_x", 0.0)
            q.center_y = s.get("center_y", 0.0)
            q.record_components = int(s.get("record_components", False))
            q.num_scattering_levels = s.get("num_scattering_levels", 0)
            q.record_statistics = int(s.get("record_statistics", False))
            q.redshift = float(s.get("redshift", 0.0))
            arr[k] = q
            self._instr.append((self._wlg[q.wavelength_grid], q.num_pixels_x * q.num_pixels_y))
        self._call("set_instruments", self._h, C.c_int32(len(instruments)), arr, C.c_int32(int(has_medium_emission)))

    def set_secondary(self, emission_grid, spatial_bias, wavelength_bias, bias_min, bias_max, temperature, planck_abs,
                      rf_sigma_abs, em_sigma_abs, rf_cmb=None):
        keep = [_d(temperature), _d(planck_abs), _d(rf_sigma_abs), _d(em_sigma_abs)]
        cmb = _d(rf_cmb) if rf_cmb is not None else (None, None)
        sec = SkSecondary(emission_grid, len(keep[0][0]), spatial_bias, wavelength_bias, bias_min, bias_max,
                          keep[0][1], keep[1][1], keep[2][1], keep[3][1], cmb[1])
        self._call("set_secondary", self._h, C.byref(sec))

    def set_secondary_media(self, emission_grid, spatial_bias, wavelength_bias, bias_min, bias_max, tables: Sequence[tuple],
                            rf_cmb=None):
        """tables: (temperature, planck_abs, rf_sigma_abs, em_sigma_abs) per dust component (sk_engine_set_secondary_media)."""
        arr = (SkSecondary * len(tables))()
        keep = []
        cmb = _d(rf_cmb) if rf_cmb is not None else (None, None)
        for h, t in enumerate(tables):
            k = [_d(x) for x in t]
            keep.append(k)
            arr[h] = SkSecondary(emission_grid, len(k[0][0]), spatial_bias, wavelength_bias, bias_min, bias_max,
                                 k[0][1], k[1][1], k[2][1], k[3][1], cmb[1])
        self._call("set_secondary_media", self._h, C.c_int32(len(tables)), arr)

    # -- running --------------------------------------------------------------------------------
    def clear_instruments(self):
        self._call("clear_instruments", self._h)

    def cuda_stream(self) -> int:
        p = C.c_void_p()
        self._call("cuda_stream", self._h, C.byref(p))
        return p.value or 0

    def device_tensor(self, which):
        """The tally buffer `which` (0 rf1, 1 rf2, 2 rf2c, 3 detector block, 4 statistics block) as a torch tensor that
        SHARES the engine's memory, for in-place collectives: a CUDA tensor for the engine (NCCL); the test-only CPU
        oracle hands out host memory (gloo)."""
        import torch
        ptr, n = self.device_buffer(which)
        if self.prefix != "sk_engine_":
            arr = np.ctypeslib.as_array(C.cast(ptr, _dp), shape=(int(n),))
            return torch.from_numpy(arr)

        class _Span:
            __cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                        "strides": None}
        return torch.as_tensor(_Span(), device=f"cuda:{self.config.device}")

    def clear_rf(self, primary=True):
        self._call("clear_rf", self._h, C.c_int32(int(primary)))

    def prepare_primary(self, num_packets):
        self._call("prepare_primary", self._h, C.c_uint64(int(num_packets)))

    def prepare_secondary(self, num_packets) -> float:
        lum = C.c_double()
        self._call("prepare_secondary", self._h, C.c_uint64(int(num_packets)), C.byref(lum))
        return lum.value

    def set_history_interleave(self, block, num_parts, part):
        """This engine runs every num_parts-th block of `block` histories of the ranges given to run_segment."""
        self._call("set_history_interleave", self._h, C.c_uint64(int(block)), C.c_uint32(int(num_parts)), C.c_uint32(int(part)))

    def run_segment(self, first, count, primary=True, peel=True, store=False, stream_id=0):
        self._call("run_segment", self._h, C.c_uint64(int(first)), C.c_uint64(int(count)), C.c_int32(int(primary)),
                   C.c_int32(int(peel)), C.c_int32(int(store)), C.c_uint32(stream_id))

    def launch_segment(self, first, count, primary=True, peel=True, store=False, stream_id=0):
        self._call("launch_segment", self._h, C.c_uint64(int(first)), C.c_uint64(int(count)),
                   C.c_int32(int(primary)), C.c_int32(int(peel)), C.c_int32(int(store)), C.c_uint32(stream_id))

    def synchronize(self):
        self._call("synchronize", self._h)

    STAGES = ("advance", "launch", "peel_setup", "detect", "sample", "trace_forward", "trace_interaction", "trace_peel")

    def last_stage_ms(self) -> dict:
        out = (C.c_float * len(self.STAGES))()
        self._call("last_stage_ms", self._h, out)
        return dict(zip(self.STAGES, [float(x) for x in out]))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._call("last_kernel_ms", self._h, C.byref(ms))
        return ms.value

    def communicate_rf(self, primary=True):
        self._call("communicate_rf", self._h, C.c_int32(int(primary)))

    def absorbed_luminosity(self, primary=True) -> float:
        out = C.c_double()
        self._call("absorbed_luminosity", self._h, C.c_int32(int(primary)), C.byref(out))
        return out.value

    # -- outputs --------------------------------------------------------------------------------
    def read_rf(self, which=0, out=None):
        """[cell][bin] radiation field table; `out` may be a caller-owned C-contiguous float64 array of that shape (page-locked:
        the engine then copies straight into it)."""
        if out is None:
            out = np.empty((self.num_cells, self.num_rf), dtype=np.float64)
        assert out.shape == (self.num_cells, self.num_rf) and out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
        self._call("read_rf", self._h, C.c_int32(which), out.ctypes.data_as(_dp))
        return out

    def read_sed(self, instrument=0, component=SK_COMP_TOTAL):
        nl, _ = self._instr[instrument]
        out = np.empty(nl, dtype=np.float64)
        self._call("read_sed", self._h, C.c_int32(instrument), C.c_int32(component), out.ctypes.data_as(_dp))
        return out

    def read_ifu(self, instrument=0, component=SK_COMP_TOTAL, out=None):
        """[ell][pixel] tallies; `out` may be a caller-owned C-contiguous float64 array of that shape (the way the C++ shim
        reads straight into FluxRecorder's arrays)."""
        nl, npix = self._instr[instrument]
        if out is None:
            out = np.empty((nl, npix), dtype=np.float64)
        assert out.shape == (nl, npix) and out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
        self._call("read_ifu", self._h, C.c_int32(instrument), C.c_int32(component), out.ctypes.data_as(_dp))
        return out

    def read_sed_stats(self, instrument=0):
        nl, _ = self._instr[instrument]
        out = np.empty((5, nl), dtype=np.float64)
        for k in range(5):
            row = np.empty(nl, dtype=np.float64)
            self._call("read_sed_stats", self._h, C.c_int32(instrument), C.c_int32(k), row.ctypes.data_as(_dp))
            out[k] = row
        return out

    def read_ifu_stats(self, instrument=0):
        """Sum w^k per frame pixel, k = 0..4: [k][ell][pixel] (FluxRecorder::_wifu)."""
        nl, npix = self._instr[instrument]
        out = np.empty((5, nl, npix), dtype=np.float64)
        for k in range(5):
            self._call("read_ifu_stats", self._h, C.c_int32(instrument), C.c_int32(k), out[k].ctypes.data_as(_dp))
        return out

    def counters(self, reset=False) -> dict:
        c = SkCounters()
        self._call("counters", self._h, C.byref(c), C.c_int32(int(reset)))
        return c.as_dict()

    def measure_gather_peak(self, num_records) -> float:
        """records/s of dependent scattered 32-byte record fetches on this device (the crossing loop's access pattern)."""
        out = C.c_double()
        self._call("measure_gather_peak", self._h, C.c_int32(int(num_records)), C.byref(out))
        return out.value

    def device_buffer(self, which):
        ptr = C.c_void_p()
        n = C.c_uint64()
        self._call("device_buffer", self._h, C.c_int32(which), C.byref(ptr), C.byref(n))
        return ptr.value, n.value
