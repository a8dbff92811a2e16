"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/ctypes front-end of oracle/edt_oracle.c with the same call signature as the
reference's Python API (src/edt.pyx:121-310, 622-734 of the reference), so parity
tests read like the reference's own tests.  Only tests/, __graft_entry__.smoke()
and bench.py's CPU-baseline leg may import this module.

`load_reference()` returns the real compiled reference (oracle/_ref, built by
oracle/Makefile from /root/reference without copying sources) or None.
"""
import ctypes
import importlib.util
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(quiet=True):
  """Compile liboracle_edt.so (and oracle/_ref when /root/reference exists)."""
  cmd = ["make", "-C", _HERE, "all"]
  subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL if quiet else None)


def _lib():
  global _LIB
  if _LIB is None:
    path = os.path.join(_HERE, "liboracle_edt.so")
    if not os.path.exists(path):
      subprocess.run(["make", "-C", _HERE, "liboracle_edt.so"], check=True,
                     stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(path)
    i64, f32, vp = ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
    lib.oracle_edtsq.argtypes = [vp, ctypes.c_int, ctypes.c_int, i64, i64, i64,
                                 f32, f32, f32, ctypes.c_int, vp]
    lib.oracle_edtsq.restype = ctypes.c_int
    lib.oracle_bruteforce_edtsq.argtypes = lib.oracle_edtsq.argtypes
    lib.oracle_bruteforce_edtsq.restype = ctypes.c_int
    lib.oracle_pass_first.argtypes = [vp, ctypes.c_int, i64, i64, i64, f32, ctypes.c_int, vp]
    lib.oracle_pass_first.restype = ctypes.c_int
    lib.oracle_pass_later.argtypes = [vp, ctypes.c_int, ctypes.c_int, i64, i64, i64, f32, ctypes.c_int,
                                      ctypes.c_int, vp]
    lib.oracle_pass_later.restype = ctypes.c_int
    _LIB = lib
  return _LIB


def load_reference():
  """The unmodified reference as a module object, or None if oracle/_ref is absent."""
  refdir = os.path.join(_HERE, "_ref")
  if not os.path.isdir(refdir):
    return None
  for name in sorted(os.listdir(refdir)):
    if name.startswith("edt.") and name.endswith(".so"):
      spec = importlib.util.spec_from_file_location("edt", os.path.join(refdir, name))
      mod = importlib.util.module_from_spec(spec)
      saved = sys.modules.get("edt")
      sys.modules["edt"] = mod       # Cython modules look themselves up by name
      try:
        spec.loader.exec_module(mod)
      finally:
        if saved is not None:
          sys.modules["edt"] = saved
        else:
          sys.modules.pop("edt", None)
      return mod
  return None


def _as_unsigned(data):
  """dtype handling of edt.pyx:670-732: ints reinterpreted as unsigned, bool as
  uint8, float32/float64 compared by value (-0.0 == +0.0)."""
  dt = data.dtype
  if dt == np.bool_:
    return data.view(np.uint8)
  if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
    return data.view(np.dtype("u%d" % dt.itemsize))
  if dt == np.float32:
    return (data + np.float32(0)).view(np.uint32)
  if dt == np.float64:
    return (data + np.float64(0)).view(np.uint64)
  return None


def _layout(data, anisotropy):
  """x-fastest dims and weights, edt.pyx:651-664 (3-D), 429-440 (2-D)."""
  nd = data.ndim
  if nd == 1:
    aniso = [float(anisotropy)] if np.isscalar(anisotropy) else [float(anisotropy[0])]
  else:
    aniso = [float(a) for a in anisotropy]
  shape = list(data.shape)
  if data.flags.f_contiguous:
    order = "F"
  else:
    order = "C"
    shape = shape[::-1]
    aniso = aniso[::-1]
  while len(shape) < 3:
    shape.append(1)
    aniso.append(1.0)
  return order, shape, aniso


def _run(fn, data, anisotropy, black_border):
  if isinstance(data, list):
    data = np.array(data)
  nd = data.ndim
  if nd > 3:
    raise TypeError("Multi-Label EDT library only supports up to 3 dimensions got {}.".format(nd))
  if data.size == 0:
    return np.zeros(shape=data.shape, dtype=np.float32)
  if not data.flags.c_contiguous and not data.flags.f_contiguous:
    data = np.ascontiguousarray(data)
  if anisotropy is None:
    anisotropy = 1.0 if nd == 1 else (1.0,) * nd
  order, (sx, sy, sz), (wx, wy, wz) = _layout(data, anisotropy)
  out = np.zeros(data.size, dtype=np.float32)
  lab = _as_unsigned(data)
  if lab is None:               # unsupported dtype: the reference returns zeros
    return out.reshape(data.shape, order=order)
  lab = np.ascontiguousarray(lab.ravel(order="K"))
  rc = fn(lab.ctypes.data, lab.dtype.itemsize, nd, sx, sy, sz,
          wx, wy, wz, int(bool(black_border)), out.ctypes.data)
  if rc != 0:
    raise RuntimeError("oracle failed")
  return out.reshape(data.shape, order=order)


def voxel_graph_edtsq(data, voxel_graph, anisotropy=None, black_border=False):
  """Restatement of _edt2dsq_voxel_graph / _edt3dsq_voxel_graph (edt_voxel_graph.hpp:54-214) as
  bound by edt.pyx:514-620 and 736-844: draw the foreground (label > 0, labels not told apart)
  on a grid of twice the resolution -- even cells are the voxels, the cell after a voxel along
  +x / +y / +z is foreground only if graph bit 0 / 2 / 4 allows that step, the other cells of
  the 2x2(x2) block are foreground, and with a black border the last cell layer of every axis is
  background -- take the binary transform at half the anisotropy, keep the even cells."""
  data = np.asarray(data)
  nd = data.ndim
  if nd not in (2, 3):
    raise TypeError("Voxel connectivity graph is only supported for 2D and 3D. Got {}.".format(nd))
  if data.size == 0:
    return np.zeros(shape=data.shape, dtype=np.float32)
  if not data.flags.c_contiguous and not data.flags.f_contiguous:
    data = np.ascontiguousarray(data)
  f_order = data.flags.f_contiguous
  graph = np.asarray(voxel_graph)
  graph = np.asfortranarray(graph) if f_order else np.ascontiguousarray(graph)
  graph = graph.view(np.uint8) if graph.dtype in (np.uint8, np.int8) else graph.astype(np.uint8)
  if anisotropy is None:
    anisotropy = (1.0,) * nd
  weights = [float(np.float32(a)) for a in anisotropy]
  if data.dtype.kind == "f":
    fg = data > 0
  elif data.dtype == np.bool_:
    fg = data
  else:
    fg = data != 0                       # astype(unsigned) > 0
  # index the arrays as [x, y(, z)] with x the memory-fastest axis (edt.pyx:429-440, 651-664)
  if not f_order:
    fg, graph, weights = fg.T, graph.T, weights[::-1]
  cells = np.zeros(tuple(2 * s for s in fg.shape), dtype=np.uint8, order="F")
  even, odd = slice(0, None, 2), slice(1, None, 2)
  edge_bits = (0x01, 0x04, 0x10)
  for offs in np.ndindex(*(2,) * nd):
    value = fg
    if sum(offs) == 1:                   # the edge cell towards the next voxel of one axis
      value = fg & ((graph & edge_bits[offs.index(1)]) != 0)
    cells[tuple(odd if o else even for o in offs)] = value
  if black_border:
    for axis in range(nd):
      last = [slice(None)] * nd
      last[axis] = -1
      cells[tuple(last)] = 0
  doubled = edtsq(cells, tuple(w / 2 for w in weights), black_border)
  out = np.asfortranarray(doubled[(even,) * nd])
  return out if f_order else np.ascontiguousarray(out.T)


def edtsq(data, anisotropy=None, black_border=False, parallel=1, voxel_graph=None, order=None):
  if voxel_graph is not None:
    return voxel_graph_edtsq(data, voxel_graph, anisotropy, black_border)
  return _run(_lib().oracle_edtsq, data, anisotropy, black_border)


def edt(data, anisotropy=None, black_border=False, parallel=1, voxel_graph=None, order=None):
  dt = edtsq(data, anisotropy, black_border, voxel_graph=voxel_graph)
  return np.sqrt(dt, dt)


def sdf(data, anisotropy=None, black_border=False, parallel=1, voxel_graph=None, order=None):
  """edt.pyx:121-158: edt(data) - edt(data == 0)."""
  data = np.asarray(data)
  dt = edt(data, anisotropy, black_border, voxel_graph=voxel_graph)
  dt -= edt(data == 0, anisotropy, black_border, voxel_graph=voxel_graph)
  return dt


def sdfsq(data, anisotropy=None, black_border=False, parallel=1, voxel_graph=None):
  """edt.pyx:161-202."""
  data = np.asarray(data)
  return (edtsq(data, anisotropy, black_border, voxel_graph=voxel_graph)
          - edtsq(data == 0, anisotropy, black_border, voxel_graph=voxel_graph))


def bruteforce_edtsq(data, anisotropy=None, black_border=False):
  """O(N^2) evaluation of the definition; tiny volumes only."""
  return _run(_lib().oracle_bruteforce_edtsq, data, anisotropy, black_border)


def pass_first(labels_zyx, wx, black_border):
  """First-axis pass of a C-ordered (z, y, x) integer volume -> float32 (z, y, x)."""
  lab = np.ascontiguousarray(labels_zyx)
  sz, sy, sx = lab.shape
  out = np.zeros(lab.shape, dtype=np.float32)
  rc = _lib().oracle_pass_first(lab.ctypes.data, lab.dtype.itemsize, sx, sy, sz, float(wx),
                                int(bool(black_border)), out.ctypes.data)
  assert rc == 0
  return out


def pass_later(labels_zyx, f_zyx, axis, w, border_lo, border_hi):
  """One envelope pass (axis 1 = y, 2 = z) of C-ordered (z, y, x) arrays, in place on f_zyx."""
  lab = np.ascontiguousarray(labels_zyx)
  assert f_zyx.flags.c_contiguous and f_zyx.dtype == np.float32
  sz, sy, sx = lab.shape
  rc = _lib().oracle_pass_later(lab.ctypes.data, lab.dtype.itemsize, int(axis), sx, sy, sz, float(w),
                                int(bool(border_lo)), int(bool(border_hi)), f_zyx.ctypes.data)
  assert rc == 0
  return f_zyx
