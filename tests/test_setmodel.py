"""tools/setmodel.py restates the CPython set behaviour that decides, in trimesh's graph.traversals, which component
is started next and at which node (the device code in csrc/shb_kernels.cu, shb_pyset_traversal_order, restates the
same rules).  Here the model is held against real sets."""
import random
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
import setmodel


def test_model_equals_real_sets():
    rnd = random.Random(20261018)
    for _ in range(600):
        n = rnd.choice([6, 12, 30, 60, 100, 143, 198, 250, 306, 307, 400, 800, 1300])
        c = rnd.choice([1, 2, 2, 3, 3, 4, 6, 10, 30])
        w = [rnd.random() ** 3 + 0.01 for _ in range(c)]
        comp = rnd.choices(range(c), weights=w, k=n)
        order = list(range(n)) * 2
        rnd.shuffle(order)
        assert setmodel.pop_order(n, comp) == setmodel.real(n, comp, order)


def test_pop_is_not_simply_the_smallest_id():
    """198 nodes, a first component of 186: difference_update rebuilds the table with 64 slots, id 67 lands in slot 3 and
    is popped before id 20 (the case that seeded random solids found on the device)."""
    comp = [0] * 198
    rest = [20, 48, 54, 58, 67, 80, 90, 100, 120, 150, 170, 190]
    for i in rest:
        comp[i] = 1
    assert setmodel.pop_order(198, comp) == [0, 67] == setmodel.real(198, comp, list(range(198)))
