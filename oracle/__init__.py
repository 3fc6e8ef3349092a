"""CPU oracle for the multiplane slicing hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this package, and only as the checker or as the timed CPU
baseline.  Nothing under ``shoulder_b200/`` imports it; the product path has no CPU fallback.

PARITY UNPINNED.  The arithmetic of this path lives in the third-party dependency *trimesh*
(required ``>=4.0.0`` by reference ``pyproject.toml:12``, locked at 3.23.5 by
``poetry.lock:4145-4146``), with shapely 2.1.0 / GEOS, networkx 2.8.8 and scipy 1.15.2
underneath.  None of them is vendored under ``/root/reference``; trimesh and shapely are not
installable here (no network), and the reference's own ``tests/`` hold no assertions, golden
vectors or known-answer files for this path.  ``trimesh_path.py`` therefore
restates trimesh's *published* algorithm (``intersections.mesh_multiplane / mesh_plane /
plane_lines``, ``path.exchange.misc.lines_to_path / edges_to_path``, ``graph.traversals /
fill_traversals / split_traversal``, ``grouping.hashable_rows / float_to_int``,
``path.Path2D.{discrete,polygons_closed,area,bounds,centroid}``) and anchors on the
reference's call sites: ``src/shoulder/humerus/slice.py:21-29`` (the sweep),
``:34-60`` (centroid / area rules), ``:65-80,166-189`` (outline choice + arc-length
resample), ``:85-147,191-206`` (centring, polar, sort / roll), ``:157-164`` (cutoff window)
and ``src/shoulder/humerus/canal.py:40-85``.  The post-trimesh part (``slice_arrays.py``)
follows the reference source, which *is* available.  Independent known answers come from
analytic solids (tests/test_oracle_analytic.py).

scipy's real ``csgraph.depth_first_order`` is used for the traversal, so that part is the
library trimesh itself calls rather than a recollection of it.
"""
from .trimesh_path import (  # noqa: F401
    TOL_MERGE,
    TOL_ZERO,
    OraclePath2D,
    mesh_plane,
    mesh_multiplane,
    section_multiplane,
    section,
    OraclePath3D,
    hashable_rows,
    float_to_int,
    rank_key,
)
from .slice_arrays import OracleSlices, radial_image, cutoff_window  # noqa: F401
