"""Imports modules of the reference UNCHANGED from /root/reference (this container only — the
GPU box has no copy, tests that need it skip there and use the committed vectors instead).

The reference's hot path needs trimesh / shapely / skspatial / plotly / onnxruntime / ruptures /
circle_fit / lsq-ellipse, none installable here.  Only what those imports NAME is stubbed; every
line of the reference module that runs is the reference's own:

  * ``shoulder.humerus.slice``           needs ``shoulder.humerus.mesh`` for a type annotation only
  * ``shoulder.humerus.canal``           needs ``skspatial.objects.Line / Points`` (best_fit = mean + first right
                                          singular vector, what scikit-spatial's Line.best_fit computes) and plotly
  * ``shoulder.humerus.bicipital_groove`` needs onnxruntime (stub raises when a session is opened; the feature
                                          stage before the classifier runs for real on scipy / sklearn)
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src/shoulder")


def available() -> bool:
    return (REF / "humerus" / "slice.py").exists()


class _Line:
    """scikit-spatial ``Line``: ``best_fit(points)`` -> point = centroid, direction = first right singular vector."""

    def __init__(self, point, direction):
        self.point, self.direction = np.asarray(point, dtype=float), np.asarray(direction, dtype=float)

    @classmethod
    def best_fit(cls, points, **kw):
        pts = np.asarray(points, dtype=float)
        c = pts.mean(axis=0)
        _, _, vh = np.linalg.svd(pts - c, **kw)
        return cls(c, vh[0])


class _OnnxStub(types.ModuleType):
    class InferenceSession:
        def __init__(self, *a, **k):
            raise RuntimeError("onnxruntime is not installed in this image")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(modname: str, path: Path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def reference_modules():
    """dict of the reference modules loaded from source: base, utils, slice, canal, bicipital_groove."""
    if _cache:
        return _cache
    if not available():
        raise FileNotFoundError(REF)
    stubs = []
    pkg = _stub("shoulder"); pkg.__path__ = [str(REF)]
    hum = _stub("shoulder.humerus"); hum.__path__ = [str(REF / "humerus")]
    pkg.humerus = hum
    meshmod = _stub("shoulder.humerus.mesh", Obb=object, FullObb=object, ProxObb=object)
    hum.mesh = meshmod
    # third-party names the landmark modules import at module level
    if "plotly" not in sys.modules:
        go = _stub("plotly.graph_objects", Scatter3d=lambda **k: k, Mesh3d=lambda **k: k, Surface=object, Figure=object)
        _stub("plotly", graph_objects=go)
        stubs += ["plotly", "plotly.graph_objects"]
    if "skspatial" not in sys.modules:
        so = _stub("skspatial.objects", Line=_Line, Points=lambda a: np.asarray(a), Plane=object, Vector=object, Sphere=object)
        _stub("skspatial", objects=so)
        stubs += ["skspatial", "skspatial.objects"]
    if "onnxruntime" not in sys.modules:
        sys.modules["onnxruntime"] = _OnnxStub("onnxruntime")
        stubs.append("onnxruntime")
    if "matplotlib" not in sys.modules:
        _stub("matplotlib.pyplot")
        _stub("matplotlib", pyplot=sys.modules["matplotlib.pyplot"])
        stubs += ["matplotlib", "matplotlib.pyplot"]
    if "trimesh" not in sys.modules:       # base.py: a type annotation (``mesh: trimesh.Trimesh``), nothing is called
        _stub("trimesh", Trimesh=object)
        stubs.append("trimesh")
    try:
        utils = _load("shoulder.utils", REF / "utils.py")
    except Exception:                      # utils imports more third-party names; the slice pin does not need it
        utils = None
    pkg.utils = utils
    try:
        base = _load("shoulder.base", REF / "base.py")
    except Exception:
        base = None
    pkg.base = base
    _cache["slice"] = _load("shoulder.humerus.slice", REF / "humerus" / "slice.py")
    hum.slice = _cache["slice"]
    for name in ("canal", "bicipital_groove"):
        try:
            _cache[name] = _load(f"shoulder.humerus.{name}", REF / "humerus" / f"{name}.py")
            setattr(hum, name, _cache[name])
        except Exception as e:             # pragma: no cover - reported by the test that needs it
            _cache[name] = e
    _cache["utils"], _cache["base"] = utils, base
    for k in stubs:                        # the stubs served the imports; nothing else should see them
        sys.modules.pop(k, None)
    return _cache


def load_extra(name: str):
    """Further landmark modules of the reference, loaded on demand (their third-party imports are stubbed as above)."""
    mods = reference_modules()
    if name in mods:
        return mods[name]
    stubs = []
    def need(modname, **attrs):
        if modname not in sys.modules:
            _stub(modname, **attrs)
            stubs.append(modname)
    need("plotly.graph_objects", Scatter3d=lambda **k: k, Mesh3d=lambda **k: k, Surface=object, Figure=object)
    need("plotly", graph_objects=sys.modules["plotly.graph_objects"])
    need("skspatial.objects", Line=_Line, Points=lambda a: np.asarray(a), Plane=object, Vector=object, Sphere=object)
    need("skspatial", objects=sys.modules["skspatial.objects"])
    if "onnxruntime" not in sys.modules:
        sys.modules["onnxruntime"] = _OnnxStub("onnxruntime"); stubs.append("onnxruntime")
    need("ellipse", LsqEllipse=object)
    need("trimesh.geometry")
    need("trimesh", Trimesh=object, geometry=sys.modules["trimesh.geometry"])
    try:
        mods[name] = _load(f"shoulder.humerus.{name}", REF / "humerus" / f"{name}.py")
    finally:
        for k in stubs:
            sys.modules.pop(k, None)
    return mods[name]


class OracleMesh:
    """What ``Slices`` touches of ``obb.mesh``: ``bounds`` and ``section_multiplane`` — the latter answered by the
    oracle's restatement of trimesh (``oracle.section_multiplane``)."""

    def __init__(self, vertices, faces, merge="hash", factory=None):
        self.vertices, self.faces = np.asarray(vertices, dtype=np.float64), np.asarray(faces, dtype=np.int64)
        self.bounds = np.array([self.vertices.min(axis=0), self.vertices.max(axis=0)])
        self._merge = merge
        self._factory = factory

    def section_multiplane(self, plane_origin, plane_normal, heights):
        if self._factory is not None:
            return self._factory(self, plane_origin, plane_normal, heights)
        import oracle
        return oracle.section_multiplane(self.vertices, self.faces, plane_origin, plane_normal, heights, merge=self._merge)


class OracleObb:
    def __init__(self, vertices, faces, transform=None, **kw):
        self.mesh = OracleMesh(vertices, faces, **kw)
        z = self.mesh.bounds[:, 2]
        self.z_bounds = (z[0], z[1])
        self.z_length = abs(z[0]) + abs(z[1])
        self.transform = np.eye(4) if transform is None else transform
        self.cutoff_pcts = [0.5, 0.8]


class Neck:
    def __init__(self, neck_z):
        self.neck_z = float(neck_z)


def exact_frame(vertices, faces, transform):
    """Frame change with elementwise numpy arithmetic only (no BLAS), so that a stored 4x4 gives bit-identical
    vertices on every host — the committed vectors depend on the exact input coordinates."""
    v = np.asarray(vertices, dtype=np.float64)
    m = np.asarray(transform, dtype=np.float64)
    out = np.empty_like(v)
    for r in range(3):
        out[:, r] = ((v[:, 0] * m[r, 0] + v[:, 1] * m[r, 1]) + v[:, 2] * m[r, 2]) + m[r, 3]
    f = np.asarray(faces, dtype=np.int64)
    if np.linalg.det(m[:3, :3]) < 0:
        f = np.ascontiguousarray(f[:, ::-1])
    return out, f
