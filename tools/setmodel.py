import random, sys
LINEAR_PROBES, PERTURB_SHIFT, MINSIZE = 9, 5, 8
def table_size_for(n):
    size, fill = MINSIZE, 0
    for _ in range(n):
        fill += 1
        if fill * 5 >= (size - 1) * 3:
            minused = fill * 2 if fill > 50000 else fill * 4
            ns = MINSIZE
            while ns <= minused: ns <<= 1
            size = ns
    return size
def insert_clean(tab, mask, key):
    perturb = key; i = key & mask
    while True:
        if tab[i] is None: tab[i] = key; return
        if i + LINEAR_PROBES <= mask:
            for j in range(1, LINEAR_PROBES + 1):
                if tab[i + j] is None: tab[i + j] = key; return
        perturb >>= PERTURB_SHIFT
        i = (i * 5 + 1 + perturb) & mask
def pop_order(n, comp):
    """comp[id] = component label; returns the ids popped as traversal starts."""
    size = table_size_for(n); mask = size - 1
    tab = [None] * size
    for i in range(n): tab[i] = i                 # final table of the construction: slot == id
    DUMMY = -1
    fill = used = n; finger = 0; starts = []
    members = {}
    for i, c in enumerate(comp): members.setdefault(c, []).append(i)
    def find(key):                                 # lookup for discard (same probe sequence as set_lookkey)
        perturb = key; i = key & mask
        while True:
            if tab[i] == key: return i
            if tab[i] is None: return None
            if i + LINEAR_PROBES <= mask:
                for j in range(1, LINEAR_PROBES + 1):
                    if tab[i + j] == key: return i + j
                    if tab[i + j] is None: return None
            perturb >>= PERTURB_SHIFT
            i = (i * 5 + 1 + perturb) & mask
    while used > 0:
        i = finger & mask
        while tab[i] is None or tab[i] == DUMMY:
            i += 1
            if i > mask: i = 0
        key = tab[i]; tab[i] = DUMMY; used -= 1; finger = i + 1
        starts.append(key)
        for k in members[comp[key]]:
            if k == key: continue
            j = find(k); tab[j] = DUMMY; used -= 1
        if fill - used > mask // 4:
            minused = used * 2 if used > 50000 else used * 4
            ns = MINSIZE
            while ns <= minused: ns <<= 1
            old = tab; tab = [None] * ns; mask = ns - 1
            for e in old:
                if e is not None and e != DUMMY: insert_clean(tab, mask, e)
            fill = used
    return starts
def real(n, comp, order):
    s = set(order)
    members = {}
    for i, c in enumerate(comp): members.setdefault(c, []).append(i)
    starts = []
    while s:
        k = s.pop(); starts.append(k)
        s.difference_update(members[comp[k]])
    return starts
if __name__ == "__main__":
    rnd = random.Random(1)
    bad = 0
    for trial in range(3000):
        n = rnd.choice([6, 12, 30, 60, 100, 143, 198, 250, 306, 307, 400, 800, 1300])
        C = rnd.choice([1, 2, 2, 3, 3, 4, 6, 10, 30])
        # skewed component sizes
        w = [rnd.random() ** 3 + 0.01 for _ in range(C)]
        comp = rnd.choices(range(C), weights=w, k=n)
        order = list(range(n)) * 2; rnd.shuffle(order)
        a, b = pop_order(n, comp), real(n, comp, order)
        if a != b:
            bad += 1
            if bad < 5: print("MISMATCH n", n, "C", C, a[:6], b[:6])
    print("mismatches", bad, "of 3000")


def pop_order_slots(n, comp):
    """Same rule, restated the way the device code works (shb_pyset_warp): no table, only the current slot of every node
    outside the first component and, during a rebuild, one bit per slot of the new table; a pop is a minimum search
    over (slot - finger) mod size."""
    members = {}
    for i, c in enumerate(comp): members.setdefault(c, []).append(i)
    c0 = comp[0]
    rid = [i for i in range(n) if comp[i] != c0]          # ascending ids of the remaining nodes
    alive = [True] * len(rid)
    rslot = list(rid)                                     # slot == id in the table of the construction
    mask = table_size_for(n) - 1
    fill, used, finger = n, n - len(members[c0]), 1
    starts = [0]
    def rebuild():
        nonlocal mask, fill
        if fill - used <= mask // 4: return
        order = sorted((k for k in range(len(rid)) if alive[k]), key=lambda k: rslot[k])
        minused = used * 2 if used > 50000 else used * 4
        ns = MINSIZE
        while ns <= minused: ns <<= 1
        mask = ns - 1
        taken = set()
        for k in order:
            key = rid[k]; perturb = key; i = key & mask; found = None
            while found is None:
                ncand = 10 if i + LINEAR_PROBES <= mask else 1
                for j in range(ncand):
                    if i + j not in taken: found = i + j; break
                if found is None:
                    perturb >>= PERTURB_SHIFT
                    i = (i * 5 + 1 + perturb) & mask
            taken.add(found); rslot[k] = found
        fill = used
    rebuild()
    while used > 0:
        f = finger & mask
        k = min((k for k in range(len(rid)) if alive[k]), key=lambda k: (rslot[k] - f) & mask)
        starts.append(rid[k]); finger = rslot[k] + 1
        c = comp[rid[k]]
        for j in range(len(rid)):
            if comp[rid[j]] == c: alive[j] = False
        used -= len(members[c])
        rebuild()
    return starts
