"""refshim -- run the reference's UNMODIFIED Python sources on a numpy stand-in for Mitsuba 3 / Dr.Jit.

TEST INFRASTRUCTURE (part of oracle/): used by tests/golden/make_refshim_golden.py to generate
golden vectors in THIS container and by a `-m "not gpu"` test that replays a small case when
/root/reference is present.  It never runs on the GPU box and nothing in the product imports it.

Why: the reference's hot path is python/integrators/volpathsimple.py, a Python file written against
the Mitsuba 3 / Dr.Jit API.  Neither library can be imported or built here, but the
reference FILE can: this module installs two stand-in modules named `mitsuba` and `drjit` that
provide exactly the API surface that file (plus python/batched.py, python/opt_config.py and
python/util.py) touches, as masked numpy lane arrays with a small reverse-mode tape, and then
imports the reference files from /root/reference WITHOUT modifying them.  What is pinned by the
golden vectors made this way:

  * everything volpathsimple.py itself decides: the path state machine, every mask, the RNG draw
    order (which sampler, which lanes, how many draws), the reservoir, the recursion through
    `sample_recursive`, the gradient formulae handed to `dr.backward_from`, the sign conventions
    of path replay -- i.e. the part of the oracle that is a *restatement of the reference*;
  * batched.py's drivers (`sample_batch_pixels`, `sample_batch_rays`, `render_batch_primal`,
    `render_batch_backward`): sub-seeds, film, dL = grad/spp, primal-then-adjoint sequencing.

What is NOT pinned (still "defined by us"): the arithmetic of the un-vendored
Mitsuba branch.  The stand-in obtains it from `uivr_oracle_shim_*` (oracle/uivr_oracle.h), thin
wrappers over the very functions the oracle's own path code calls, or restates small public
formulae (PCG32, TEA, `mis_weight`) in numpy.

The reference restarts the supergrid DDA at every tentative collision and re-derives positions from
accumulated distances (volpathsimple.py:331-334, :365-367, :497-499); the oracle carries one DDA
along the segment.  The two agree to float32 rounding, not bit for bit, so goldens made here are
compared with a stated tolerance.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import importlib.util
import os
import sys
import types
from typing import List

import numpy as np

from . import oracle as O

F32 = np.float32
F64 = np.float64
INV_4PI = F32(1.0 / (4.0 * np.pi))
LARGEST = float(np.finfo(np.float32).max)
REF_ROOT = "/root/reference"


# ======================================================================================
# native primitives (the "upstream" arithmetic)
# ======================================================================================

_sigs_done = False


def _lib():
    global _sigs_done
    L = O.lib()
    if not _sigs_done:
        fp, dp, u8p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        i32p, u32p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        vp = C.c_void_p
        L.uivr_oracle_shim_create.argtypes = [C.POINTER(O._Scene), fp, fp]
        L.uivr_oracle_shim_create.restype = vp
        L.uivr_oracle_shim_destroy.argtypes = [vp]
        L.uivr_oracle_shim_film_uv.argtypes = [vp, C.c_int, u32p, fp, fp, fp, fp]
        L.uivr_oracle_shim_camera_ray.argtypes = [vp, C.c_int, fp, i32p, fp, fp, fp, fp]
        L.uivr_oracle_shim_box_entry.argtypes = [C.c_int, fp, fp, fp, i32p]
        L.uivr_oracle_shim_entry_spawn.argtypes = [C.c_int, fp, fp, fp, fp]
        L.uivr_oracle_shim_exit.argtypes = [C.c_int, fp, fp, fp, u8p]
        L.uivr_oracle_shim_dir_to_local.argtypes = [vp, C.c_int, fp, fp]
        L.uivr_oracle_shim_sample_interaction.argtypes = [vp, C.c_int, fp, fp, fp, fp, u8p, fp, fp, fp, u8p]
        L.uivr_oracle_shim_sample_interaction_drt.argtypes = [vp, C.c_int, fp, fp, fp, u64p, u64p, u8p, fp, fp, fp, u8p]
        L.uivr_oracle_shim_lookup.argtypes = [vp, C.c_int, C.c_int, fp, fp]
        L.uivr_oracle_shim_scatter.argtypes = [vp, C.c_int, C.c_int, fp, fp, u8p, dp]
        L.uivr_oracle_shim_uniform_sphere.argtypes = [C.c_int, fp, fp, fp]
        L.uivr_oracle_shim_fma.argtypes = [C.c_int, fp, fp, fp, fp]
        L.uivr_oracle_shim_env_eval.argtypes = [vp, C.c_int, fp, fp, fp]
        L.uivr_oracle_shim_env_sample.argtypes = [vp, C.c_int, fp, fp, fp, fp, fp]
        L.uivr_oracle_shim_env_eval.restype = L.uivr_oracle_shim_env_sample.restype = None
        for f in ("destroy", "film_uv", "camera_ray", "box_entry", "entry_spawn", "exit", "dir_to_local",
                  "sample_interaction", "sample_interaction_drt", "lookup", "scatter", "uniform_sphere", "fma"):
            getattr(L, "uivr_oracle_shim_" + f).restype = None
        _sigs_done = True
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _c32(a, n):
    """contiguous float32 array of n lanes (broadcast width-1 inputs)"""
    a = np.asarray(a, dtype=F32)
    if a.shape[0] != n:
        a = np.broadcast_to(a, (n,) + a.shape[1:])
    return np.ascontiguousarray(a)


def _fma(a, b, c):
    n = max(a.shape[0], b.shape[0], c.shape[0])
    a, b, c = _c32(a, n), _c32(b, n), _c32(c, n)
    out = np.empty(n, dtype=F32)
    _lib().uivr_oracle_shim_fma(n, _p(a, C.c_float), _p(b, C.c_float), _p(c, C.c_float), _p(out, C.c_float))
    return out


class _Session:
    """One scene (medium grids + sensor + emitter) bound to a native shim handle."""
    current: "_Session" = None

    def __init__(self, desc, sigma_t, albedo):
        self.desc = desc
        x, y, z = desc["res"]
        self.sigma_t = np.ascontiguousarray(np.asarray(sigma_t, dtype=F32).reshape(z, y, x))
        self.albedo = np.ascontiguousarray(np.asarray(albedo, dtype=F32).reshape(z, y, x, 3))
        self.sc = O.make_scene(desc, dict(max_depth=1))
        self.h = _lib().uivr_oracle_shim_create(C.byref(self.sc), _p(self.sigma_t, C.c_float), _p(self.albedo, C.c_float))
        assert self.h
        self.dsigma = np.zeros((z, y, x, 1), dtype=F64)
        self.dalbedo = np.zeros((z, y, x, 3), dtype=F64)
        self.radiance = np.asarray(desc["radiance"], dtype=F32)
        self.envmap = desc.get("env_data") is not None
        self.draws = 0

    def close(self):
        _lib().uivr_oracle_shim_destroy(self.h)
        self.h = None

    def __enter__(self):
        self._prev = _Session.current
        _Session.current = self
        self._err = np.seterr(all="ignore")
        return self

    def __exit__(self, *a):
        np.seterr(**self._err)
        _Session.current = self._prev
        self.close()


def _S() -> _Session:
    assert _Session.current is not None, "refshim: no active scene session"
    return _Session.current


# ======================================================================================
# tape (reverse mode, just enough for dr.backward_from on the expressions of the reference)
# ======================================================================================

class _AD:
    enabled = False


class _Node:
    __slots__ = ("parents", "sink")

    def __init__(self, parents=None, sink=None):
        self.parents = parents
        self.sink = sink


def _backprop(node, g):
    if node.sink is not None:
        node.sink(g)
        return
    for p, w in node.parents:
        _backprop(p, g * w)


# ======================================================================================
# lane arrays
# ======================================================================================

class Float:
    __slots__ = ("v", "node")

    def __init__(self, v=0.0):
        self.node = None
        if isinstance(v, Float):
            self.v = v.v.copy()
            self.node = v.node
        elif isinstance(v, _Int):
            self.v = v.v.astype(F32)
        else:
            self.v = np.array(v, dtype=F32, ndmin=1)

    @staticmethod
    def _zeros(n=1):
        return Float(np.zeros(n, dtype=F32))

    def __len__(self):
        return self.v.shape[0]

    def __repr__(self):
        return f"Float({self.v})"

    # -- element access
    def __getitem__(self, k):
        if isinstance(k, Mask):
            return Float(self)
        return float(self.v[k if self.v.shape[0] > 1 else 0])

    def __setitem__(self, k, val):
        r = select(k, val, self)
        self.v, self.node = r.v, r.node

    def _assign(self, r):
        self.v, self.node = r.v, r.node
        return self

    # -- arithmetic
    def __add__(self, o): return _bin(self, o, np.add, lambda a, b: 1.0, lambda a, b: 1.0)
    def __radd__(self, o): return _bin(o, self, np.add, lambda a, b: 1.0, lambda a, b: 1.0)
    def __sub__(self, o): return _bin(self, o, np.subtract, lambda a, b: 1.0, lambda a, b: -1.0)
    def __rsub__(self, o): return _bin(o, self, np.subtract, lambda a, b: 1.0, lambda a, b: -1.0)
    def __mul__(self, o): return _bin(self, o, np.multiply, lambda a, b: b, lambda a, b: a)
    def __rmul__(self, o): return _bin(o, self, np.multiply, lambda a, b: b, lambda a, b: a)
    def __truediv__(self, o): return _bin(self, o, np.divide, lambda a, b: 1.0 / b, lambda a, b: -a / (b * b))
    def __rtruediv__(self, o): return _bin(o, self, np.divide, lambda a, b: 1.0 / b, lambda a, b: -a / (b * b))
    def __neg__(self): return _un(self, np.negative, lambda a: -1.0)
    def __iadd__(self, o): return self._assign(self + o)
    def __isub__(self, o): return self._assign(self - o)
    def __imul__(self, o): return self._assign(self * o)
    def __itruediv__(self, o): return self._assign(self / o)

    # -- comparisons
    def __lt__(self, o): return Mask(self.v < _F(o).v)
    def __le__(self, o): return Mask(self.v <= _F(o).v)
    def __gt__(self, o): return Mask(self.v > _F(o).v)
    def __ge__(self, o): return Mask(self.v >= _F(o).v)


def _raw(v, node=None):
    r = Float.__new__(Float)
    r.v = v
    r.node = node
    return r


def _F(x) -> Float:
    return x if isinstance(x, Float) else Float(x)


def _bin(a, b, fn, da, db):
    if isinstance(a, (Vec, Struct)) or isinstance(b, (Vec, Struct)):
        return NotImplemented
    a, b = _F(a), _F(b)
    r = _raw(fn(a.v, b.v))
    if _AD.enabled and (a.node is not None or b.node is not None):
        av, bv = a.v.astype(F64), b.v.astype(F64)
        ps = []
        if a.node is not None:
            ps.append((a.node, da(av, bv)))
        if b.node is not None:
            ps.append((b.node, db(av, bv)))
        r.node = _Node(ps)
    return r


def _un(a, fn, da):
    a = _F(a)
    r = _raw(fn(a.v))
    if _AD.enabled and a.node is not None:
        r.node = _Node([(a.node, da(a.v.astype(F64)))])
    return r


class Mask:
    __slots__ = ("v",)

    def __init__(self, v=False):
        self.v = v.v.copy() if isinstance(v, Mask) else np.array(v, dtype=bool, ndmin=1)

    @staticmethod
    def _zeros(n=1):
        return Mask(np.zeros(n, dtype=bool))

    def __len__(self):
        return self.v.shape[0]

    def __repr__(self):
        return f"Mask({self.v})"

    def __getitem__(self, k):
        return Mask(self) if isinstance(k, Mask) else bool(self.v[k])

    def __setitem__(self, k, val):
        self.v = np.where(_M(k).v, _M(val).v, self.v)

    def __and__(self, o):
        if isinstance(o, (MaskVec, Vec)):
            return NotImplemented
        return Mask(self.v & _M(o).v)

    def __or__(self, o):
        if isinstance(o, MaskVec):
            return NotImplemented
        return Mask(self.v | _M(o).v)

    __rand__ = __and__
    __ror__ = __or__

    def __xor__(self, o): return Mask(self.v ^ _M(o).v)
    def __invert__(self): return Mask(~self.v)

    def __iand__(self, o):
        self.v = self.v & _M(o).v
        return self

    def __ior__(self, o):
        self.v = self.v | _M(o).v
        return self


def _M(x) -> Mask:
    return x if isinstance(x, Mask) else Mask(x)


class _Int:
    __slots__ = ("v",)
    DT = np.int32

    def __init__(self, v=0):
        if isinstance(v, _Int):
            self.v = v.v.astype(self.DT)
        elif isinstance(v, Float):
            self.v = v.v.astype(np.int64).astype(self.DT)  # truncation, as a C cast
        else:
            self.v = np.array(v, dtype=np.int64, ndmin=1).astype(self.DT)

    @classmethod
    def _zeros(cls, n=1):
        return cls(np.zeros(n, dtype=cls.DT))

    def __len__(self):
        return self.v.shape[0]

    def __repr__(self):
        return f"{type(self).__name__}({self.v})"

    def _o(self, o):
        return o.v if isinstance(o, _Int) else np.asarray(o, dtype=np.int64)

    def __getitem__(self, k):
        return type(self)(self) if isinstance(k, Mask) else int(self.v[k])

    def __setitem__(self, k, val):
        self.v = np.where(_M(k).v, type(self)(val).v, self.v).astype(self.DT)

    def __add__(self, o):
        if isinstance(o, (Float, float)):
            return Float(self) + o
        return type(self)((self.v.astype(np.int64) + self._o(o)))
    __radd__ = __add__
    def __sub__(self, o):
        if isinstance(o, (Float, float)):
            return Float(self) - o
        return type(self)((self.v.astype(np.int64) - self._o(o)))
    def __mul__(self, o):
        if isinstance(o, (Float, float)):
            return Float(self) * o
        return type(self)((self.v.astype(np.int64) * self._o(o)))
    def __rmul__(self, o):
        if isinstance(o, (Float, float)):
            return o * Float(self)
        return type(self)((self.v.astype(np.int64) * self._o(o)))
    def __floordiv__(self, o): return type(self)((self.v.astype(np.int64) // self._o(o)))
    def __lt__(self, o): return Mask(self.v.astype(np.int64) < self._o(o))
    def __le__(self, o): return Mask(self.v.astype(np.int64) <= self._o(o))
    def __gt__(self, o): return Mask(self.v.astype(np.int64) > self._o(o))
    def __ge__(self, o): return Mask(self.v.astype(np.int64) >= self._o(o))


class Int32(_Int):
    __slots__ = ()
    DT = np.int32


class UInt32(_Int):
    __slots__ = ()
    DT = np.uint32


class _U64:
    """loop-state leaf for the PCG32 state"""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v


# ---- small fixed-size arrays of lane arrays ------------------------------------------

class Vec:
    N = 3
    __slots__ = ("c",)

    def __init__(self, *a):
        n = self.N
        if len(a) == 0:
            self.c = [Float(0.0) for _ in range(n)]
        elif len(a) == 1:
            x = a[0]
            if isinstance(x, (Vec, VecU)):
                assert x.N == n
                self.c = [Float(t) for t in x.c]
            elif isinstance(x, (list, tuple)) or (isinstance(x, np.ndarray) and x.ndim >= 1 and x.shape[-1] == n and not isinstance(x, Float)):
                if isinstance(x, np.ndarray) and x.ndim == 2:
                    self.c = [Float(x[:, i]) for i in range(n)]
                else:
                    self.c = [Float(x[i]) for i in range(n)]
            else:
                self.c = [Float(x) for _ in range(n)]
        else:
            assert len(a) == n
            self.c = [Float(t) for t in a]

    @classmethod
    def _zeros(cls, n=1):
        return cls(*[Float._zeros(n) for _ in range(cls.N)])

    x = property(lambda s: s.c[0], lambda s, v: s.c.__setitem__(0, Float(v)))
    y = property(lambda s: s.c[1], lambda s, v: s.c.__setitem__(1, Float(v)))
    z = property(lambda s: s.c[2], lambda s, v: s.c.__setitem__(2, Float(v)))

    def __len__(self):
        return self.N

    def __repr__(self):
        return f"{type(self).__name__}({[t.v for t in self.c]})"

    def numpy(self):
        n = max(len(t) for t in self.c)
        return np.stack([np.broadcast_to(t.v, (n,)) for t in self.c], axis=-1).astype(F32)

    def _other(self, o):
        if isinstance(o, (Vec, VecU)):
            assert o.N == self.N
            return o.c
        if isinstance(o, (list, tuple, np.ndarray)) and not isinstance(o, Float) and len(o) == self.N:
            return [o[i] for i in range(self.N)]
        return [o] * self.N

    def _map(self, o, f):
        oc = self._other(o)
        r = type(self).__new__(type(self))
        r.c = [f(a, b) for a, b in zip(self.c, oc)]
        return r

    def __add__(self, o): return self._map(o, lambda a, b: a + b)
    def __radd__(self, o): return self._map(o, lambda a, b: b + a)
    def __sub__(self, o): return self._map(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._map(o, lambda a, b: b - a)
    def __mul__(self, o): return self._map(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._map(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._map(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._map(o, lambda a, b: b / a)

    def __neg__(self):
        r = type(self).__new__(type(self))
        r.c = [-a for a in self.c]
        return r

    def _assign(self, r):
        for a, b in zip(self.c, r.c):
            a._assign(b)
        return self

    def __iadd__(self, o): return self._assign(self + o)
    def __isub__(self, o): return self._assign(self - o)
    def __imul__(self, o): return self._assign(self * o)
    def __itruediv__(self, o): return self._assign(self / o)

    def __and__(self, m):
        mc = m.c if isinstance(m, MaskVec) else [m] * self.N
        r = type(self).__new__(type(self))
        r.c = [select(k, a, 0.0) for a, k in zip(self.c, mc)]
        return r

    __rand__ = __and__

    def __iand__(self, m): return self._assign(self & m)

    def __getitem__(self, k):
        if isinstance(k, Mask):
            return type(self)(self)
        return self.c[k]

    def __setitem__(self, k, val):
        if isinstance(k, Mask):
            for a, b in zip(self.c, self._other(val)):
                a[k] = b
        else:
            self.c[k] = Float(val)

    def _cmp(self, o, f):
        return MaskVec([f(a, b) for a, b in zip(self.c, self._other(o))])

    def __gt__(self, o): return self._cmp(o, lambda a, b: a > b)
    def __ge__(self, o): return self._cmp(o, lambda a, b: a >= b)
    def __lt__(self, o): return self._cmp(o, lambda a, b: a < b)
    def __le__(self, o): return self._cmp(o, lambda a, b: a <= b)


class Vector3f(Vec):
    __slots__ = ()


class Point3f(Vec):
    __slots__ = ()


class Color3f(Vec):
    __slots__ = ()


class Vector4f(Vec):
    N = 4
    __slots__ = ()


class Vec2(Vec):
    N = 2
    __slots__ = ()


class Point2f(Vec2):
    __slots__ = ()


class Vector2f(Vec2):
    __slots__ = ()


class VecU:
    """2-vector of UInt32 lane arrays (mi.Point2u)"""
    N = 2

    def __init__(self, x, y=None):
        if isinstance(x, (Vec, VecU)):
            self.c = [UInt32(t) for t in x.c]
        else:
            self.c = [UInt32(x), UInt32(y)]

    x = property(lambda s: s.c[0])
    y = property(lambda s: s.c[1])


class ScalarVector2(tuple):
    """mi.ScalarVector2u / 2f: a plain host-side pair (float32-rounded values held as Python floats)"""

    def __new__(cls, *a):
        v = a[0] if len(a) == 1 else a
        return tuple.__new__(cls, (float(F32(v[0])), float(F32(v[1]))))

    def __eq__(self, o):
        return (self[0] == o[0], self[1] == o[1])

    __hash__ = tuple.__hash__

    def __mul__(self, o):
        return o.__rmul__(self) if isinstance(o, Vec) else NotImplemented

    def rcp(self):
        return ScalarVector2(float(F32(1.0) / F32(self[0])), float(F32(1.0) / F32(self[1])))


class MaskVec:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [_M(m) for m in c]

    N = property(lambda s: len(s.c))

    def __and__(self, o):
        oc = o.c if isinstance(o, MaskVec) else [o] * len(self.c)
        return MaskVec([a & b for a, b in zip(self.c, oc)])

    __rand__ = __and__

    def __or__(self, o):
        oc = o.c if isinstance(o, MaskVec) else [o] * len(self.c)
        return MaskVec([a | b for a, b in zip(self.c, oc)])

    __ror__ = __or__

    def __invert__(self):
        return MaskVec([~a for a in self.c])


# ---- records --------------------------------------------------------------------------

class Struct:
    FIELDS: dict = {}
    ZERO: dict = {}

    def __init__(self, other=None):
        if other is not None:
            for f, T in self.FIELDS.items():
                object.__setattr__(self, f, _copy_field(T, getattr(other, f, None)))
            self._copy_extra(other)
        else:
            for f, T in self.FIELDS.items():
                object.__setattr__(self, f, T(self.ZERO[f]) if f in self.ZERO else T())

    def _copy_extra(self, other):
        pass

    @classmethod
    def _zeros(cls, n=1):
        r = cls()
        for f, T in cls.FIELDS.items():
            v = T._zeros(n)
            if f in cls.ZERO:
                v = v + cls.ZERO[f] if not isinstance(v, Mask) else v
            object.__setattr__(r, f, v)
        return r

    def __setattr__(self, name, value):
        T = self.FIELDS.get(name)
        if T is not None:
            value = _copy_field(T, value)
        object.__setattr__(self, name, value)

    def __getitem__(self, k):
        assert isinstance(k, Mask)
        return type(self)(self)

    def __setitem__(self, k, other):
        assert isinstance(k, Mask)
        for f in self.FIELDS:
            src = getattr(other, f, None)
            if src is not None:
                getattr(self, f)[k] = src


def _copy_field(T, v):
    if v is None:
        return T()
    return T(v)


class Ray3f(Struct):
    # `mode`: 0 = origin not known to be inside the medium, 1 = inside, 2 = has left the medium
    FIELDS = {"o": Point3f, "d": Vector3f, "maxt": Float, "time": Float, "mode": Int32}
    ZERO = {"maxt": LARGEST}
    wavelengths = None

    def __init__(self, other=None, d=None):
        if d is not None:  # Ray3f(o, d)
            Struct.__init__(self)
            self.o, self.d = other, d
        else:
            Struct.__init__(self, other)

    def __call__(self, t):
        t = _F(t)
        r = Point3f.__new__(Point3f)
        r.c = [_raw(_fma(t.v, self.d.c[a].v, self.o.c[a].v)) for a in range(3)]
        return r


RayDifferential3f = Ray3f


class Interaction3f(Struct):
    FIELDS = {"t": Float, "p": Point3f}
    ZERO = {"t": np.inf}

    def is_valid(self):
        return neq(self.t, np.inf)

    def spawn_ray(self, d):
        r = Ray3f()
        r.o, r.d = self.p, d
        r.maxt = Float(LARGEST)
        r.mode = Int32(np.ones(max(width(self.p), width(d)), dtype=np.int32))
        return r


class SurfaceInteraction3f(Interaction3f):
    # kind: 0 = none, 1 = entry into the medium box, 2 = exit
    FIELDS = {"t": Float, "p": Point3f, "kind": Int32, "ro": Point3f, "rd": Vector3f}
    ZERO = {"t": np.inf}

    def spawn_ray(self, d):
        """cross the null boundary: into the medium from an entry hit, out of it from an exit hit"""
        n = max(width(self.t), width(d))
        ro, rd, t = _c32(self.ro.numpy(), n), _c32(self.rd.numpy(), n), _c32(self.t.v, n)
        o_in = np.empty((n, 3), dtype=F32)
        _lib().uivr_oracle_shim_entry_spawn(n, _p(ro, C.c_float), _p(rd, C.c_float), _p(t, C.c_float), _p(o_in, C.c_float))
        entry = np.broadcast_to(self.kind.v == 1, (n,))
        r = Ray3f()
        r.o = Point3f(np.where(entry[:, None], o_in, np.broadcast_to(self.p.numpy(), (n, 3))))
        r.d = d
        r.maxt = Float(LARGEST)
        r.mode = Int32(np.where(entry, 1, 2))
        return r

    def emitter(self, scene):
        return EmitterPtr(~self.is_valid())

    def target_medium(self, d):
        return MediumPtr(self.is_valid())


class MediumInteraction3f(Interaction3f):
    FIELDS = {"t": Float, "p": Point3f, "mint": Float, "sigma_s": Color3f, "sigma_n": Color3f,
              "sigma_t": Color3f, "combined_extinction": Color3f}
    ZERO = {"t": np.inf}
    medium = None

    def _copy_extra(self, other):
        object.__setattr__(self, "medium", getattr(other, "medium", None))

    def __setitem__(self, k, other):
        Struct.__setitem__(self, k, other)
        if getattr(other, "medium", None) is not None:
            object.__setattr__(self, "medium", other.medium)


class DirectionSample3f(Struct):
    FIELDS = {"d": Vector3f, "pdf": Float}

    def __init__(self, *a):
        Struct.__init__(self, a[0] if len(a) == 1 else None)
        if len(a) == 3:  # DirectionSample3f(scene, si, ref): direction of the ray that produced `si`
            self.d = a[1].rd


class _Ptr:
    """array of instance pointers with a single possible target: just a validity mask"""

    def __init__(self, valid=True):
        self.valid = _M(valid)

    @classmethod
    def _zeros(cls, n=1):
        return cls(np.zeros(n, dtype=bool))

    def __setitem__(self, k, other):
        self.valid[k] = other.valid


# ======================================================================================
# drjit stand-in
# ======================================================================================

def width(x) -> int:
    if isinstance(x, SensorPtr):
        return len(x.index)
    if isinstance(x, Tensor):
        return int(x.v.size)
    leaves: List = []
    _flatten(x, leaves)
    return max([len(l.v) for l in leaves], default=1)


def select(m, a, b):
    if isinstance(a, Tensor) or isinstance(b, Tensor):   # python/losses.py:18-21 (huber)
        return Tensor(np.where(m, a.v if isinstance(a, Tensor) else F32(a), b.v if isinstance(b, Tensor) else F32(b)))
    if isinstance(m, MaskVec) or isinstance(a, Vec) or isinstance(b, Vec):
        proto = a if isinstance(a, Vec) else (b if isinstance(b, Vec) else Color3f())
        n = proto.N
        mc = m.c if isinstance(m, MaskVec) else [m] * n
        ac = a.c if isinstance(a, Vec) else [a] * n
        bc = b.c if isinstance(b, Vec) else [b] * n
        r = type(proto).__new__(type(proto))
        r.c = [select(k, x, y) for k, x, y in zip(mc, ac, bc)]
        return r
    if isinstance(a, Struct):
        r = type(a)(b)
        r[_M(m)] = a
        return r
    m = _M(m)
    if isinstance(a, _Int) and isinstance(b, _Int):
        return type(a)(np.where(m.v, a.v, b.v))
    if isinstance(a, Mask) or isinstance(b, Mask):
        return Mask(np.where(m.v, _M(a).v, _M(b).v))
    a, b = _F(a), _F(b)
    r = _raw(np.where(m.v, a.v, b.v))
    if _AD.enabled and (a.node is not None or b.node is not None):
        ps = []
        if a.node is not None:
            ps.append((a.node, m.v.astype(F64)))
        if b.node is not None:
            ps.append((b.node, (~m.v).astype(F64)))
        r.node = _Node(ps)
    return r


def _lift(f):
    def g(a, *rest):
        if isinstance(a, Vec):
            r = type(a).__new__(type(a))
            r.c = [f(t, *rest) for t in a.c]
            return r
        return f(a, *rest)
    return g


rcp = _lift(lambda a: a.rcp() if isinstance(a, ScalarVector2) else 1.0 / _F(a))
sqr = _lift(lambda a: a * a if isinstance(a, Tensor) else _F(a) * _F(a))


def sqrt_t(a):
    """dr.sqrt as python/losses.py uses it: on a tensor, or on a python float"""
    return Tensor(np.sqrt(a.v)) if isinstance(a, Tensor) else float(np.sqrt(a))


def log_t(a):
    """dr.log as python/losses.py:50-51 uses it: on a tensor, or on a python float"""
    return Tensor(np.log(a.v)) if isinstance(a, Tensor) else float(np.log(a))


def _minmax(fn):
    def g(a, b):
        if isinstance(a, Tensor) or isinstance(b, Tensor):
            return Tensor(fn(a.v if isinstance(a, Tensor) else F32(a), b.v if isinstance(b, Tensor) else F32(b)))
        if isinstance(a, Vec) or isinstance(b, Vec):
            proto = a if isinstance(a, Vec) else b
            ac = a.c if isinstance(a, Vec) else [a] * proto.N
            bc = b.c if isinstance(b, Vec) else [b] * proto.N
            r = type(proto).__new__(type(proto))
            r.c = [g(x, y) for x, y in zip(ac, bc)]
            return r
        a, b = _F(a), _F(b)
        r = _raw(fn(a.v, b.v))
        if _AD.enabled and (a.node is not None or b.node is not None):
            pick_a = (r.v == a.v)
            ps = []
            if a.node is not None:
                ps.append((a.node, pick_a.astype(F64)))
            if b.node is not None:
                ps.append((b.node, (~pick_a).astype(F64)))
            r.node = _Node(ps)
        return r
    return g


minimum = _minmax(np.minimum)
maximum = _minmax(np.maximum)


def exp(a):
    return _un(a, np.exp, lambda v: np.exp(v))


def clip(v, lo, hi):
    return Tensor(np.minimum(np.maximum(v.v, F32(lo)), F32(hi))) if isinstance(v, Tensor) else minimum(maximum(v, lo), hi)


def shape(x):
    return x.shape


def prod(x):
    return int(np.prod(x)) if len(x) else 1


def hmin(x):
    return min(x) if isinstance(x, (tuple, list)) else _raw(np.minimum(np.minimum(x.c[0].v, x.c[1].v), x.c[2].v))


def ravel(x):
    if isinstance(x, Vec):
        return _raw(np.ascontiguousarray(x.numpy()).reshape(-1))
    return x.array if isinstance(x, Tensor) else x


def hsum(x):
    return Tensor(np.asarray(x.v, dtype=F64).sum().astype(F32).reshape(1)) if isinstance(x, Tensor) else x


def habs(x):
    return Tensor(np.abs(x.v)) if isinstance(x, Tensor) else _un(x, np.abs, lambda v: np.sign(v))


def hmax(a):
    return _raw(np.maximum(np.maximum(a.c[0].v, a.c[1].v), a.c[2].v))


def mean(a):
    """horizontal mean of a 3-vector, in the oracle's operation order"""
    return ((a.c[0] + a.c[1]) + a.c[2]) * F32(1.0 / 3.0)


def any_(m):
    if isinstance(m, MaskVec):
        r = m.c[0]
        for k in m.c[1:]:
            r = r | k
        return r
    if isinstance(m, Mask):
        return bool(m.v.any())
    return bool(np.any(m))


def all_(m):
    if isinstance(m, MaskVec):
        r = m.c[0]
        for k in m.c[1:]:
            r = r & k
        return r
    if isinstance(m, Mask):
        return bool(m.v.all())
    return bool(np.all(m))


def neq(a, b):
    if isinstance(a, _Ptr):
        assert b is None
        return Mask(a.valid)
    if isinstance(a, Vec):
        return a._cmp(b, lambda x, y: neq(x, y))
    if isinstance(a, _Int):
        return Mask(a.v.astype(np.int64) != a._o(b))
    if isinstance(a, Mask):
        return a ^ b
    return Mask(_F(a).v != _F(b).v)


def eq(a, b):
    r = neq(a, b)
    return ~r


def isfinite(a):
    return Mask(np.isfinite(_F(a).v))


def detach(x, preserve_type=True):
    if isinstance(x, Float):
        return _raw(x.v.copy())
    if isinstance(x, Vec):
        r = type(x).__new__(type(x))
        r.c = [detach(t) for t in x.c]
        return r
    if isinstance(x, Struct):
        r = type(x)(x)
        for f in r.FIELDS:
            object.__setattr__(r, f, detach(getattr(r, f)))
        return r
    if isinstance(x, Tensor):
        return Tensor(x)
    return x


def zeros(T, n=1):
    return T._zeros(n)


def empty(T, n=1):
    return T._zeros(n)


def full(T, value, n=1):
    return T(np.full(n, value))


def arange(T, n):
    return T(np.arange(n))


def gather(T, src, idx, active=True):
    i = idx.v.astype(np.int64)
    if isinstance(src, SensorPtr):
        return SensorPtr(src.frames, src.index[i])
    if isinstance(src, VecU):
        return VecU(UInt32(src.c[0].v[i]), UInt32(src.c[1].v[i]))
    if isinstance(src, Float) and isinstance(T, type) and issubclass(T, Vec):
        r = T.__new__(T)
        r.c = [Float(src.v[i * T.N + c]) for c in range(T.N)]  # N consecutive entries per index
        return r
    if isinstance(src, Vec):
        r = type(src).__new__(type(src))
        r.c = [Float(t.v[i]) for t in src.c]
        return r
    return type(src)(src.v[i])


@contextlib.contextmanager
def _grad_scope(value, when=True):
    old = _AD.enabled
    if when:
        _AD.enabled = value
    try:
        yield
    finally:
        _AD.enabled = old


def resume_grad(*a, when=True):
    return _grad_scope(True, when)


def suspend_grad(*a, when=True):
    return _grad_scope(False, when)


def backward_from(x):
    for t in (x.c if isinstance(x, Vec) else [x]):
        if isinstance(t, Float) and t.node is not None:
            _backprop(t.node, np.ones(t.v.shape[0], dtype=F64))


class _GradLeaf:
    def __init__(self, n):
        self.g = np.zeros(n, dtype=F64)

    def __call__(self, g):
        self.g = self.g + g


def enable_grad(x):
    for t in (x.c if isinstance(x, Vec) else [x]):
        t.node = _Node(sink=_GradLeaf(len(t)))


def grad(x):
    if isinstance(x, Vec):
        r = type(x).__new__(type(x))
        r.c = [grad(t) for t in x.c]
        return r
    return Float(x.node.sink.g.astype(F32)) if x.node is not None and x.node.sink is not None else Float(0.0)


def set_grad(x, g):
    x._grad_in = np.asarray(g.v if isinstance(g, Tensor) else g, dtype=F64)


def enqueue(mode, x):
    _AD_queue.append(x)


def traverse(T, mode):
    while _AD_queue:
        x = _AD_queue.pop()
        x._backward(x._grad_in)


_AD_queue: List = []


class ADMode:
    Primal = 0
    Forward = 1
    Backward = 2


class CustomOp:
    def grad_out(self):
        return self._grad_out

    def set_grad_out(self, v):
        self._grad_fwd = v


def custom(Op, *args):
    op = Op()
    out = op.eval(*args)
    first = out[0] if isinstance(out, tuple) else out
    first._custom_op = op
    return out


def _flatten(obj, out):
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return
    if isinstance(obj, (Float, Mask, _Int, _U64)):
        out.append(obj)
    elif isinstance(obj, (Vec, VecU, MaskVec)):
        out.extend(obj.c)
    elif isinstance(obj, Struct):
        for f in obj.FIELDS:
            _flatten(getattr(obj, f), out)
    elif isinstance(obj, _Ptr):
        out.append(obj.valid)
    elif isinstance(obj, (tuple, list)):
        for o in obj:
            _flatten(o, out)
    elif hasattr(obj, "loop_put"):
        col = _Collector()
        obj.loop_put(col)
        for fn in col.fns:
            _flatten(fn(), out)


class _Collector:
    def __init__(self):
        self.fns = []

    def put(self, fn):
        self.fns.append(fn)


class Loop(_Collector):
    """Wavefront-style masked loop: the body runs for all lanes while any lane is active; at every
    `loop(cond)` the state of the lanes that were NOT active in the iteration just executed is
    rolled back to what it was before that iteration (what Dr.Jit's loop recording guarantees)."""

    def __init__(self, name, state=None):
        _Collector.__init__(self)
        self.name = name
        if state is not None:
            self.put(state)
        self._saved = None
        self._cond = None
        self._iters = 0

    def _leaves(self):
        out: List = []
        for fn in self.fns:
            _flatten(fn(), out)
        return out

    def __call__(self, cond):
        self._iters += 1
        if self._iters > 1000000:
            raise RuntimeError(f"refshim: loop {self.name!r} does not terminate")
        leaves = self._leaves()
        if self._saved is not None:
            assert len(leaves) == len(self._saved)
            m = self._cond
            for leaf, old in zip(leaves, self._saved):
                leaf.v = np.where(m, leaf.v, old)
                if isinstance(leaf, Float):
                    leaf.node = None
        self._cond = cond.v.copy()
        self._saved = [leaf.v.copy() for leaf in leaves]
        return bool(self._cond.any())


def _make_drjit():
    m = types.ModuleType("drjit")
    m.__dict__.update(dict(
        ADMode=ADMode, width=width, select=select, rcp=rcp, sqr=sqr, minimum=minimum, maximum=maximum,
        max=hmax, mean=mean, exp=exp, clip=clip, shape=shape, prod=prod, min=hmin, ravel=ravel, sum=hsum, hsum_async=hsum, sqrt=sqrt_t, log=log_t, abs=habs, any=any_, all=all_, neq=neq, eq=eq, isfinite=isfinite, detach=detach,
        zeros=zeros, empty=empty, full=full, arange=arange, gather=gather, resume_grad=resume_grad,
        suspend_grad=suspend_grad, backward_from=backward_from, enable_grad=enable_grad, grad=grad,
        set_grad=set_grad, enqueue=enqueue, traverse=traverse, CustomOp=CustomOp, custom=custom,
        nan=float("nan"), inf=float("inf"), largest=lambda T: LARGEST, detached_t=lambda T: T,
        schedule=lambda *a: None, eval=lambda *a: None, is_llvm_v=lambda T: True,
        log2i=lambda x: int(x).bit_length() - 1,
        fma=lambda a, b, c: _raw(_fma(_F(a).v, _F(b).v, _F(c).v)),
    ))
    return m


# ======================================================================================
# mitsuba stand-in
# ======================================================================================

def tea32(v0, v1, rounds=4):
    """mi.sample_tea_32 (public TEA construction), vectorised, numpy uint32"""
    scalar = not isinstance(v0, (np.ndarray, _Int)) and not isinstance(v1, (np.ndarray, _Int))
    a = np.atleast_1d(np.asarray(v0.v if isinstance(v0, _Int) else v0, dtype=np.uint64) & 0xFFFFFFFF).astype(np.uint32)
    b = np.atleast_1d(np.asarray(v1.v if isinstance(v1, _Int) else v1, dtype=np.uint64) & 0xFFFFFFFF).astype(np.uint32)
    a, b = np.broadcast_arrays(a, b)
    a, b = a.copy(), b.copy()
    s = np.zeros(1, dtype=np.uint32)
    for _ in range(rounds):
        s = s + np.uint32(0x9e3779b9)
        a = a + ((((b << np.uint32(4)) + np.uint32(0xa341316c)) ^ (b + s)) ^ ((b >> np.uint32(5)) + np.uint32(0xc8013ea4)))
        b = b + ((((a << np.uint32(4)) + np.uint32(0xad90777d)) ^ (a + s)) ^ ((a >> np.uint32(5)) + np.uint32(0x7e95761e)))
    if scalar:
        return int(a[0]), int(b[0])
    return UInt32(a), UInt32(b)


_PCG_MULT = np.uint64(0x5851f42d4c957f2d)


class Sampler:
    """`independent` sampler: one PCG32 stream per lane, seeded PCG32(initstate, initseq) =
    TEA(seed, lane index) (public PCG32 recurrence; stream assignment per SURVEY App. B.2)."""

    def __init__(self):
        self.state = _U64(np.zeros(0, dtype=np.uint64))
        self.inc = np.zeros(0, dtype=np.uint64)
        self._spp = 4
        self.draws = 0

    def seed(self, seed, wavefront_size):
        n = int(wavefront_size)
        v0, v1 = tea32(np.full(n, int(seed) & 0xFFFFFFFF, dtype=np.uint64), np.arange(n, dtype=np.uint64))
        self.state = _U64(np.zeros(n, dtype=np.uint64))
        self.inc = (v1.v.astype(np.uint64) << np.uint64(1)) | np.uint64(1)
        self._next_u32(True)
        self.state.v = self.state.v + v0.v.astype(np.uint64)
        self._next_u32(True)
        self.draws = 0

    def _next_u32(self, active):
        m = np.broadcast_to(_M(active).v, self.state.v.shape)
        old = self.state.v
        self.state.v = np.where(m, old * _PCG_MULT + self.inc, old)
        self.draws += int(m.sum())
        xs = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)).astype(np.uint32)
        rot = (old >> np.uint64(59)).astype(np.uint32)
        return (xs >> rot) | (xs << ((np.uint32(0) - rot) & np.uint32(31)))

    def next_1d(self, active=True):
        u = self._next_u32(active)
        return _raw(((u >> np.uint32(9)) | np.uint32(0x3f800000)).view(F32) - F32(1.0))

    def next_2d(self, active=True):
        a = self.next_1d(active)
        b = self.next_1d(active)
        return Point2f(a, b)

    def clone(self):
        s = Sampler()
        s.state = _U64(self.state.v.copy())
        s.inc = self.inc.copy()
        s._spp = self._spp
        return s

    def fork(self):
        s = Sampler()
        s._spp = self._spp
        return s

    def wavefront_size(self):
        return self.state.v.shape[0]

    def sample_count(self):
        return self._spp

    def set_sample_count(self, spp):
        self._spp = int(spp)

    def set_samples_per_wavefront(self, spp):
        pass

    def loop_put(self, loop):
        loop.put(lambda: (self.state,))


class Properties(dict):
    pass


class PhaseFunctionContext:
    def __init__(self, sampler=None):
        self.sampler = sampler


def _dir_to_local(w):
    n = w.shape[0]
    d = np.empty((n, 3), dtype=F32)
    _lib().uivr_oracle_shim_dir_to_local(_S().h, n, _p(w, C.c_float), _p(d, C.c_float))
    return d


def _uniform_sphere_local(u2: Vec2, n):
    """warp::square_to_uniform_sphere followed by the world->local direction transform"""
    a, b = _c32(u2.c[0].v, n), _c32(u2.c[1].v, n)
    w = np.empty((n, 3), dtype=F32)
    _lib().uivr_oracle_shim_uniform_sphere(n, _p(a, C.c_float), _p(b, C.c_float), _p(w, C.c_float))
    return _dir_to_local(w)


class PhaseFunctionPtr(_Ptr):
    """isotropic phase function"""

    def sample(self, ctx, mei, sample1, sample2, active=True):
        m = _M(active) & self.valid
        n = max(width(sample2), len(m))
        wo = Vector3f(_uniform_sphere_local(sample2, n))
        return wo, select(m, Float(INV_4PI), 0.0)

    def eval(self, ctx, mei, wo, active=True):
        return select(_M(active) & self.valid, Float(INV_4PI), 0.0)


def _env_eval(d: Vec):
    """envmap radiance and solid-angle density for rays leaving along local direction d"""
    n = width(d)
    dd = _c32(d.numpy(), n)
    le = np.empty((n, 3), dtype=F32)
    pdf = np.empty(n, dtype=F32)
    _lib().uivr_oracle_shim_env_eval(_S().h, n, _p(dd, C.c_float), _p(le, C.c_float), _p(pdf, C.c_float))
    return Color3f(le), _raw(pdf)


class EmitterPtr(_Ptr):
    """the scene's environment emitter: `constant`, or `envmap` when the session has one"""

    def pdf_direction(self, it, ds, active=True):
        m = _M(active) & self.valid
        if _S().envmap:
            return select(m, _env_eval(ds.d)[1], 0.0)
        return select(m, Float(INV_4PI), 0.0)

    def eval(self, si, active=True):
        m = _M(active) & self.valid
        if _S().envmap:
            return select(m, _env_eval(si.rd)[0], 0.0)
        return select(m, Color3f(list(_S().radiance)), 0.0)


class _SigmaLeaf:
    def __init__(self, p, mask, which, ch=0):
        self.p, self.mask, self.which, self.ch = p, mask, which, ch

    def __call__(self, g):
        S = _S()
        n = self.p.shape[0]
        g = np.where(self.mask, np.broadcast_to(g, (n,)), 0.0)
        if self.which == 0:
            gg = np.ascontiguousarray(g, dtype=F32)
            grid = S.dsigma
        else:
            gg = np.zeros((n, 3), dtype=F32)
            gg[:, self.ch] = g
            grid = S.dalbedo
        mk = np.ascontiguousarray(self.mask & (g != 0.0) & np.isfinite(g), dtype=np.uint8)
        _lib().uivr_oracle_shim_scatter(S.h, self.which, n, _p(self.p, C.c_float), _p(gg, C.c_float),
                                        _p(mk, C.c_uint8), _p(grid, C.c_double))


class Medium:
    """heterogeneous medium: sigma_t = scale * grid, albedo grid, majorant supergrid"""

    _majorant_resolution_factor = 0

    def phase_function(self):
        return PhaseFunctionPtr(True)

    def majorant_resolution_factor(self):
        return Medium._majorant_resolution_factor

    def set_majorant_resolution_factor(self, f):
        Medium._majorant_resolution_factor = int(f)

    @staticmethod
    def _points(p: Vec, mask):
        n = max(width(p), len(mask))
        pts = _c32(p.numpy(), n).copy()
        m = np.broadcast_to(mask, (n,)).copy()
        pts[~m] = 0.5
        pts[np.isnan(pts).any(axis=1)] = 0.5
        return pts, m, n

    def _lookup(self, which, pts, m, n, attach=True):
        out = np.zeros((n, 3 if which == 1 else 1), dtype=F32)
        _lib().uivr_oracle_shim_lookup(_S().h, which, n, _p(pts, C.c_float), _p(out, C.c_float))
        out[~m] = 0.0
        if which == 1:
            r = Color3f(out)
            if _AD.enabled and attach:
                for c in range(3):
                    r.c[c].node = _Node(sink=_SigmaLeaf(pts, m, 1, c))
            return r
        f = _raw(out[:, 0].copy())
        if _AD.enabled and attach and which == 0:
            f.node = _Node(sink=_SigmaLeaf(pts, m, 0))
        return f

    def _coefficients(self, p, mask):
        pts, m, n = self._points(p, mask)
        st = self._lookup(0, pts, m, n)
        maj = self._lookup(2, pts, m, n)
        sigma_t = Color3f(st, st, st)
        sigma_n = Color3f(maj, maj, maj) - sigma_t
        return Color3f(0.0), sigma_n, sigma_t, Color3f(maj, maj, maj)

    def get_scattering_coefficients(self, mei, active=True):
        s, n, t, _ = self._coefficients(mei.p, _M(active).v)
        return s, n, t

    def get_majorant(self, mei, active=True):
        pts, m, n = self._points(mei.p, _M(active).v)
        maj = self._lookup(2, pts, m, n)
        return Color3f(maj, maj, maj)

    def get_albedo(self, mei, active=True):
        pts, m, n = self._points(mei.p, _M(active).v)
        return self._lookup(1, pts, m, n)

    def get_emission(self, mei, active=True):
        """RGB emission grid (the session's second grid), trilinear, unscaled"""
        return self.get_albedo(mei, active)

    def sample_interaction(self, ray, sample, channel, active=True):
        S = _S()
        n = max(width(ray), len(_F(sample)), len(_M(active)))
        o, d = _c32(ray.o.numpy(), n), _c32(ray.d.numpy(), n)
        maxt, u = _c32(ray.maxt.v, n), _c32(_F(sample).v, n)
        act = np.ascontiguousarray(np.broadcast_to(_M(active).v, (n,)), dtype=np.uint8)
        t = np.empty(n, dtype=F32)
        st = np.empty(n, dtype=F32)
        sb = np.empty(n, dtype=F32)
        valid = np.empty(n, dtype=np.uint8)
        _lib().uivr_oracle_shim_sample_interaction(S.h, n, _p(o, C.c_float), _p(d, C.c_float), _p(maxt, C.c_float),
                                                   _p(u, C.c_float), _p(act, C.c_uint8), _p(t, C.c_float),
                                                   _p(st, C.c_float), _p(sb, C.c_float), _p(valid, C.c_uint8))
        mei = MediumInteraction3f()
        object.__setattr__(mei, "medium", self)
        mei.t = _raw(t)
        mei.p = ray(mei.t)
        vm = valid.astype(bool)
        sig = _raw(st)
        if _AD.enabled:
            pts = mei.p.numpy().copy()
            pts[~vm] = 0.5
            sig.node = _Node(sink=_SigmaLeaf(pts, vm, 0))
        maj = _raw(sb)
        mei.sigma_t = Color3f(sig, sig, sig)
        mei.combined_extinction = Color3f(maj, maj, maj)
        mei.sigma_n = mei.combined_extinction - mei.sigma_t
        mei.sigma_s = Color3f(0.0)
        mei.mint = Float(0.0)
        return mei

    def sample_interaction_drt(self, ray, sampler, channel, active=True):
        S = _S()
        n = sampler.wavefront_size()
        o, d, maxt = _c32(ray.o.numpy(), n), _c32(ray.d.numpy(), n), _c32(ray.maxt.v, n)
        act = np.ascontiguousarray(np.broadcast_to(_M(active).v, (n,)), dtype=np.uint8)
        state = np.ascontiguousarray(sampler.state.v)
        inc = np.ascontiguousarray(sampler.inc)
        t = np.empty(n, dtype=F32)
        st = np.empty(n, dtype=F32)
        wt = np.empty(n, dtype=F32)
        valid = np.empty(n, dtype=np.uint8)
        _lib().uivr_oracle_shim_sample_interaction_drt(S.h, n, _p(o, C.c_float), _p(d, C.c_float), _p(maxt, C.c_float),
                                                       _p(state, C.c_uint64), _p(inc, C.c_uint64), _p(act, C.c_uint8),
                                                       _p(t, C.c_float), _p(st, C.c_float), _p(wt, C.c_float),
                                                       _p(valid, C.c_uint8))
        sampler.state.v = state
        mei = MediumInteraction3f()
        object.__setattr__(mei, "medium", self)
        mei.t = _raw(t)
        mei.p = ray(mei.t)
        sig = _raw(st)
        mei.sigma_t = Color3f(sig, sig, sig)
        w = _raw(wt)
        return mei, Color3f(w, w, w)


class MediumPtr(_Ptr):
    """per-lane pointer to the scene's single medium (nerf.py:70-73)"""

    def get_scattering_coefficients(self, mei, active=True):
        return Medium().get_scattering_coefficients(mei, _M(active) & self.valid)

    def get_emission(self, mei, active=True):
        return Medium().get_emission(mei, _M(active) & self.valid)


class Shape:
    def __init__(self, medium):
        self._medium = medium

    def interior_medium(self):
        return self._medium


class Scene:
    def __init__(self):
        self._medium = Medium()
        self._shapes = [Shape(self._medium)]

    def shapes(self):
        return self._shapes

    def integrator(self):
        return None

    def ray_intersect(self, ray, active=True):
        n = max(width(ray), len(_M(active)))
        o, d = _c32(ray.o.numpy(), n), _c32(ray.d.numpy(), n)
        act = np.broadcast_to(_M(active).v, (n,))
        mode = np.broadcast_to(ray.mode.v, (n,))
        t_in = np.empty(n, dtype=F32)
        kind = np.empty(n, dtype=np.int32)
        _lib().uivr_oracle_shim_box_entry(n, _p(o, C.c_float), _p(d, C.c_float), _p(t_in, C.c_float), _p(kind, C.c_int32))
        t_out = np.empty(n, dtype=F32)
        ok = np.empty(n, dtype=np.uint8)
        _lib().uivr_oracle_shim_exit(n, _p(o, C.c_float), _p(d, C.c_float), _p(t_out, C.c_float), _p(ok, C.c_uint8))
        outside, inside = act & (mode == 0), act & (mode == 1)
        t = np.full(n, np.inf, dtype=F32)
        k = np.zeros(n, dtype=np.int32)
        hit_in = outside & (kind == 1)
        far = outside & (kind == 2)
        hit_out = inside & ok.astype(bool)
        t[hit_in] = t_in[hit_in]
        k[hit_in] = 1
        t[far] = t_in[far]
        k[far] = 2
        t[hit_out] = t_out[hit_out]
        k[hit_out] = 2
        si = SurfaceInteraction3f()
        si.t = _raw(t)
        si.kind = Int32(k)
        si.ro, si.rd = ray.o, ray.d
        si.p = ray(select(Mask(np.isfinite(t)), si.t, 0.0))
        return si

    def sample_emitter_direction(self, ref, sample, test_visibility=True, active=True):
        m = _M(active)
        n = max(width(sample), len(m))
        ds = DirectionSample3f()
        if _S().envmap:
            a, b = _c32(sample.c[0].v, n), _c32(sample.c[1].v, n)
            d = np.empty((n, 3), dtype=F32)
            pdf = np.empty(n, dtype=F32)
            le = np.empty((n, 3), dtype=F32)
            _lib().uivr_oracle_shim_env_sample(_S().h, n, _p(a, C.c_float), _p(b, C.c_float), _p(d, C.c_float),
                                               _p(pdf, C.c_float), _p(le, C.c_float))
            ds.d = Vector3f(d)
            ds.pdf = select(m, _raw(pdf), 0.0)
            ok = m & (ds.pdf > 0.0)
            return ds, select(ok, Color3f(le) / ds.pdf, 0.0)
        ds.d = Vector3f(_uniform_sphere_local(sample, n))
        ds.pdf = select(m, Float(INV_4PI), 0.0)
        val = select(m, Color3f(list(_S().radiance)) / Float(INV_4PI), 0.0)
        return ds, val


class SensorPtr:
    """array of perspective sensors (16-float frames, uivr_oracle_batch layout)"""

    def __init__(self, frames, index=None):
        self.frames = np.ascontiguousarray(frames, dtype=F32).reshape(-1, 16)
        self.index = np.arange(self.frames.shape[0]) if index is None else np.asarray(index)
        self.v = self.index  # width() support

    def __len__(self):
        return len(self.index)

    def sample_ray_differential(self, time, sample1, sample2, sample3, active=True):
        n = len(self.index)
        u, v = _c32(sample2.c[0].v, n), _c32(sample2.c[1].v, n)
        idx = np.ascontiguousarray(self.index, dtype=np.int32)
        return _camera_rays(self.frames, idx, u, v), Color3f(1.0)


def _camera_rays(frames, idx, u, v):
    n = u.shape[0]
    o = np.empty((n, 3), dtype=F32)
    d = np.empty((n, 3), dtype=F32)
    _lib().uivr_oracle_shim_camera_ray(_S().h, n, None if frames is None else _p(frames, C.c_float),
                                       None if idx is None else _p(idx, C.c_int32), _p(u, C.c_float),
                                       _p(v, C.c_float), _p(o, C.c_float), _p(d, C.c_float))
    r = Ray3f()
    r.o, r.d = Point3f(o), Vector3f(d)
    r.maxt = Float(LARGEST)
    r.mode = Int32(np.zeros(n, dtype=np.int32))
    return r


# ---- film (box filter, weight channel), only what batched.py drives -------------------

class Tensor:
    """mi.TensorXf: an n-d float32 array (host numpy here)"""

    def __init__(self, v, shape=None):
        if isinstance(v, Tensor):
            v = v.v
        elif isinstance(v, Float):
            v = v.v
        self.v = np.array(v, dtype=F32)
        if shape is not None:
            self.v = self.v.reshape(tuple(int(x) for x in shape))

    shape = property(lambda s: tuple(int(x) for x in s.v.shape))
    array = property(lambda s: _raw(np.ascontiguousarray(s.v).reshape(-1)))

    def numpy(self):
        return self.v

    def _bin(self, o, f):
        return Tensor(f(self.v, o.v if isinstance(o, Tensor) else o))

    def __sub__(self, o): return self._bin(o, np.subtract)
    def __add__(self, o): return self._bin(o, np.add)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __radd__(self, o): return Tensor(np.add(F32(o), self.v))
    def __rsub__(self, o): return Tensor(np.subtract(F32(o), self.v))
    def __rmul__(self, o): return Tensor(np.multiply(F32(o), self.v))
    def __lt__(self, o): return self.v < (o.v if isinstance(o, Tensor) else F32(o))   # a plain bool array: only dr.select reads it


class ImageBlock:
    def __init__(self, w, h):
        self.w, self.h = w, h
        self.entries = []
        self._coalesce = True

    def coalesce(self):
        return self._coalesce

    def set_coalesce(self, v):
        self._coalesce = bool(v)

    def channel_count(self):
        return 5

    def tensor(self):
        return None

    def put(self, pos, wavelengths=None, value=None, alpha=None, weight=1.0, active=True):
        self.entries.append((pos, value, _F(weight)))


class Film:
    def __init__(self, d):
        self.w, self.h = int(d["width"]), int(d["height"])
        assert d.get("rfilter", {}).get("type", "box") == "box"
        self.block = None

    def prepare(self, aovs):
        assert not aovs

    def flags(self):
        return 0

    def create_block(self):
        return ImageBlock(self.w, self.h)

    def put_block(self, block):
        self.block = block

    def crop_size(self):
        return ScalarVector2(self.w, self.h)

    def rfilter(self):
        return types.SimpleNamespace(is_box_filter=lambda: True)

    def sample_border(self):
        return False

    def develop(self):
        """hdrfilm: colour channels / weight channel; the box filter puts a sample into floor(pos)"""
        w, h = self.w, self.h
        acc = np.zeros((h * w, 3), dtype=F64)
        wsum = np.zeros(h * w, dtype=F64)
        pix_of = []
        for pos, value, weight in self.block.entries:
            n = max(width(pos), width(value))
            px = np.floor(np.broadcast_to(pos.c[0].v, (n,))).astype(np.int64)
            py = np.floor(np.broadcast_to(pos.c[1].v, (n,))).astype(np.int64)
            pix = py * w + px
            np.add.at(acc, pix, value.numpy().astype(F64) * np.broadcast_to(weight.v, (n,))[:, None])
            np.add.at(wsum, pix, np.broadcast_to(weight.v, (n,)).astype(F64))
            pix_of.append(pix)
        img = Tensor((acc / np.maximum(wsum, 1e-30)[:, None]).astype(F32).reshape(h, w, 3))
        entries = self.block.entries

        def backward(g):
            g = np.asarray(g, dtype=F64).reshape(h * w, -1)[:, :3]
            for (pos, value, weight), pix in zip(entries, pix_of):
                for c in range(3):
                    if value.c[c].node is not None:
                        _backprop(value.c[c].node, g[pix, c] * weight.v / wsum[pix])
        img._backward = backward
        return img


class _Bitmap:
    class PixelFormat:
        RGB = "rgb"
        RGBA = "rgba"
        XYZ = "xyz"


class _FilmFlags:
    Special = 4


_INTEGRATORS = {}


def register_integrator(name, factory):
    _INTEGRATORS[name] = factory


def load_dict(d):
    t = d["type"]
    if t == "independent":
        return Sampler()
    if t == "hdrfilm":
        return Film(d)
    if t in _INTEGRATORS:
        return _INTEGRATORS[t](Properties({k: v for k, v in d.items() if k != "type"}))
    raise NotImplementedError(f"refshim: mi.load_dict type {t!r}")


class Integrator:
    pass


class RBIntegrator(Integrator):
    """mi.ad.integrators.common.RBIntegrator: only the property handling volpathsimple.py inherits"""

    def __init__(self, props=None):
        props = props or {}
        md = props.get("max_depth", 6)
        if md < 0 and md != -1:
            raise Exception('"max_depth" must be set to -1 (infinite) or a value >= 0')
        self.max_depth = md if md != -1 else 0xFFFFFFFF
        self.rr_depth = props.get("rr_depth", 5)
        if self.rr_depth <= 0:
            raise Exception('"rr_depth" must be set to a value greater than zero!')

    def aovs(self):
        return []


def mis_weight(pdf_a, pdf_b):
    """mi.ad.common.mis_weight: power heuristic"""
    pdf_a, pdf_b = _F(pdf_a), _F(pdf_b)
    a2 = pdf_a * pdf_a
    w = a2 / _raw(_fma(pdf_b.v, pdf_b.v, a2.v))
    return detach(select(pdf_a > 0.0, w, 0.0))


def _make_mitsuba():
    m = types.ModuleType("mitsuba")
    ns = types.SimpleNamespace
    m.__dict__.update(dict(
        Float=Float, Int32=Int32, UInt32=UInt32, Mask=Mask, Bool=Mask, Spectrum=Color3f, Color3f=Color3f,
        Vector3f=Vector3f, Vector4f=Vector4f, Point3f=Point3f, Point2f=Point2f, Vector2f=Vector2f, Point2u=VecU,
        ScalarVector2u=ScalarVector2, ScalarVector2f=ScalarVector2,
        Ray3f=Ray3f, RayDifferential3f=Ray3f, Interaction3f=Interaction3f,
        SurfaceInteraction3f=SurfaceInteraction3f, MediumInteraction3f=MediumInteraction3f,
        DirectionSample3f=DirectionSample3f, PhaseFunctionContext=PhaseFunctionContext,
        PhaseFunctionPtr=PhaseFunctionPtr, EmitterPtr=EmitterPtr, SensorPtr=SensorPtr, MediumPtr=MediumPtr,
        Properties=Properties, Loop=Loop, Scene=Scene, Sampler=Sampler, Integrator=Integrator,
        SamplingIntegrator=Integrator, Film=Film, Bitmap=_Bitmap, FilmFlags=_FilmFlags, TensorXf=Tensor,
        SceneParameters=dict, has_flag=lambda flags, f: bool(flags & f), is_spectral=False,
        sample_tea_32=tea32, register_integrator=register_integrator, load_dict=load_dict,
        ad=ns(common=ns(mis_weight=mis_weight, _ReparamWrapper=None),
              integrators=ns(common=ns(RBIntegrator=RBIntegrator)),
              Adam=None, SGD=None),
    ))
    return m


# ======================================================================================
# loading the reference, drivers
# ======================================================================================

_ref = None


def available(ref_root: str = REF_ROOT) -> bool:
    return os.path.exists(os.path.join(ref_root, "python", "integrators", "volpathsimple.py"))


def load_reference(ref_root: str = REF_ROOT):
    """Import the reference's own files (unmodified, from where they lie) against the stand-ins."""
    global _ref
    if _ref is not None:
        return _ref
    if not available(ref_root):
        raise FileNotFoundError(f"reference sources not found under {ref_root}")
    names = ("drjit", "mitsuba", "util", "losses", "opt_config", "batched", "optimize")
    saved = {k: sys.modules.pop(k, None) for k in names}
    sys.modules["drjit"] = _make_drjit()
    sys.modules["mitsuba"] = _make_mitsuba()
    pydir = os.path.join(ref_root, "python")
    sys.path.insert(0, pydir)
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # /root/reference is read-only
    try:
        mods = {}
        for name, rel in (("volpathsimple", "integrators/volpathsimple.py"), ("nerf", "integrators/nerf.py"),
                          ("batched", "batched.py"), ("opt_config", "opt_config.py"), ("optimize", "optimize.py"),
                          ("losses", "losses.py")):
            spec = importlib.util.spec_from_file_location("refshim_ref_" + name, os.path.join(pydir, rel))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        sys.dont_write_bytecode = dont
        sys.path.remove(pydir)
        for k in names:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    _ref = types.SimpleNamespace(**mods)
    return _ref


def make_integrator(config_name: str, max_depth: int, **overrides):
    """Through the reference's own registry: opt_config.get_int_config(name).create(max_depth=...)
    (python/opt_config.py:83-169) -> mi.load_dict -> the factory volpathsimple.py registered."""
    ref = load_reference()
    cfg = ref.opt_config.get_int_config(config_name)
    return cfg.create(max_depth=max_depth, **overrides)


def _primary(desc, seed, spp):
    """RBIntegrator.render's preamble (upstream; restated in batched.py:134-173): seed the path
    sampler for the whole wavefront, pixel = index // spp, jitter = sampler.next_2d()."""
    S = _S()
    n = int(desc["width"]) * int(desc["height"]) * int(spp)
    sampler = Sampler()
    sampler.set_sample_count(spp)
    sampler.seed(seed, n)
    pix = (np.arange(n, dtype=np.uint32) // np.uint32(spp)).astype(np.uint32)
    jit = sampler.next_2d()
    u = np.empty(n, dtype=F32)
    v = np.empty(n, dtype=F32)
    _lib().uivr_oracle_shim_film_uv(S.h, n, _p(pix, C.c_uint32), _p(_c32(jit.c[0].v, n), C.c_float),
                                    _p(_c32(jit.c[1].v, n), C.c_float), _p(u, C.c_float), _p(v, C.c_float))
    return sampler, pix, _camera_rays(None, None, u, v)


def _film_mean(L, pix, npix, spp):
    acc = np.zeros((npix, 3), dtype=F64)
    np.add.at(acc, pix.astype(np.int64), L.astype(F64))
    return (acc / spp).astype(F32)


def render_forward(desc, integrator, sigma_t, albedo, seed, spp):
    """mi.render(scene, sensor, integrator, seed, spp) primal -> (image (H,W,3), per-sample L (S,3))."""
    h, w = int(desc["height"]), int(desc["width"])
    with _Session(desc, sigma_t, albedo):
        sampler, pix, rays = _primary(desc, seed, spp)
        with suspend_grad():
            L, valid, _ = integrator.sample(mode=ADMode.Primal, scene=Scene(), sampler=sampler, ray=rays,
                                            depth=UInt32(0), δL=None, state_in=None, reparam=None,
                                            active=Mask(True))
        Ls = _c32(L.numpy(), len(pix))
        return _film_mean(Ls, pix, h * w, spp).reshape(h, w, 3), Ls


def render_backward(desc, integrator, sigma_t, albedo, grad_image, seed_grad, spp_grad):
    """RBIntegrator.render_backward (upstream; restated in batched.py:212-326): primal pass on a CLONE
    of the sampler, dL = grad_image[pixel] / spp (box film), adjoint pass on the sampler itself.
    -> (d sigma_t (Z,Y,X,1) f64, d albedo (Z,Y,X,3) f64, per-sample primal L (S,3))."""
    h, w = int(desc["height"]), int(desc["width"])
    with _Session(desc, sigma_t, albedo) as S:
        sampler, pix, rays = _primary(desc, seed_grad, spp_grad)
        scene = Scene()
        with suspend_grad():
            L, valid, state = integrator.sample(mode=ADMode.Primal, scene=scene, sampler=sampler.clone(), ray=rays,
                                                δL=None, state_in=None, active=Mask(True), reparam=None)
            g = np.asarray(grad_image, dtype=F32).reshape(h * w, 3)[pix.astype(np.int64)] * F32(1.0 / spp_grad)
            integrator.sample(mode=ADMode.Backward, scene=scene, sampler=sampler, ray=rays, δL=Color3f(g),
                              state_in=state, active=Mask(True), reparam=None)
        return S.dsigma.copy(), S.dalbedo.copy(), _c32(L.numpy(), len(pix))


def render_batch(desc, integrator, sigma_t, albedo, sensors16, film_size, batch_size, seed, spp, spp_grad=0,
                 seed_grad=0, grad_image_fn=None):
    """python/batched.py `render_batch` (+ its backward through `_BatchedRenderOp`), run unmodified.
    -> dict(image (B,3), sensor_idx, pixels[, dsigma, dalbedo]); grad_image_fn(image) -> d loss / d image."""
    ref = load_reference()
    with _Session(desc, sigma_t, albedo) as S:
        sensors = SensorPtr(sensors16)
        out = ref.batched.render_batch(batch_size, Scene(), sensors, ScalarVector2(*film_size), params=None,
                                       integrator=integrator, film=None, pixel_format=None, sampler=None,
                                       seed=seed, seed_grad=seed_grad, spp=spp, spp_grad=spp_grad)
        image, film, render_sampler, sensor_idx, pixels = out
        res = dict(image=image.v.reshape(batch_size, 3).copy(), sensor_idx=sensor_idx.v.copy(),
                   pixels=np.stack([pixels.c[0].v, pixels.c[1].v], axis=-1))
        if grad_image_fn is not None:
            op = image._custom_op
            op._grad_out = (Tensor(np.asarray(grad_image_fn(res["image"]), dtype=F32).reshape(1, batch_size, 3)),)
            op.backward()
            res["dsigma"], res["dalbedo"] = S.dsigma.copy(), S.dalbedo.copy()
        return res
