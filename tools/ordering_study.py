"""CPU study behind the reduced-system ordering of bslam_finalize (pyslam_b200/csrc/solver.cu).

The tile Cholesky is bound by the chain of dependent diagonal tiles, i.e. by the HEIGHT of the elimination
tree of the supernode graph.  This script builds the supernode graphs of BASELINE configs 2 and 4, applies
the candidate orders (table order, nested dissection as first written, nested dissection with the balanced
tie-break, minimum degree, min-fill, shallowest-subtree-first greedy) and prints tree height and fill.
Output of the round-1 run: profiles/r1_ordering_study.md."""
import sys
import numpy as np
sys.path.insert(0, '.')
from pyslam_b200 import synthetic


def height_and_fill(n, adj, order):
    pos = {u: i for i, u in enumerate(order)}
    m = np.zeros((n, n), bool)
    for u in range(n):
        for v in adj[u]:
            a, b = pos[u], pos[v]
            m[max(a, b), min(a, b)] = True
    h, fill = [1] * n, 0
    for k in range(n):
        rows = [i for i in range(k + 1, n) if m[i, k]]
        fill += len(rows)
        for a in rows:
            h[a] = max(h[a], h[k] + 1)
            for b in rows:
                if a > b:
                    m[a, b] = True
    return max(h), fill


def nested_dissection(n, adj, balanced):
    order, side = [], [0] * n
    todo = [(list(range(n)), False)]
    while todo:
        nodes, emit = todo.pop()
        if emit or len(nodes) <= 2:
            order += nodes
            continue
        mid = len(nodes) // 2
        left, right = nodes[:mid], nodes[mid:]
        for u in left: side[u] = 1
        for u in right: side[u] = 2
        sepL = [u for u in left if any(side[v] == 2 for v in adj[u])]
        sepR = [u for u in right if any(side[v] == 1 for v in adj[u])]
        for u in nodes: side[u] = 0
        if balanced and len(sepL) == len(sepR):
            useL = max(len(left) - len(sepL), len(right)) <= max(len(left), len(right) - len(sepR))
        else:
            useL = len(sepL) <= len(sepR)
        sep = sepL if useL else sepR
        if len(sep) * 2 >= len(nodes):
            order += nodes
            continue
        ss = set(sep)
        rest = [u for u in (left if useL else right) if u not in ss]
        todo.append((sep, True))
        todo.append((right if useL else rest, False))
        todo.append((rest if useL else left, False))
    return order


def greedy(n, adj, key):
    nb = [set(a) for a in adj]; alive = [True] * n; order = []; h = [1] * n
    for _ in range(n):
        u = min((i for i in range(n) if alive[i]), key=lambda i: key(nb, h, i))
        order.append(u); alive[u] = False
        ns = list(nb[u])
        for a in ns:
            nb[a].discard(u); h[a] = max(h[a], h[u] + 1)
        for a in ns:
            for b in ns:
                if a != b: nb[a].add(b)
    return order


def fill_of(nb, i):
    ns = list(nb[i])
    return sum(1 for x in range(len(ns)) for y in range(x + 1, len(ns)) if ns[y] not in nb[ns[x]])


def study(name, n, adj):
    print('### %s (%d supernodes, mean degree %.1f)\n' % (name, n, np.mean([len(a) for a in adj])))
    print('| order | elimination-tree height | fill (off-diagonal tiles) |\n|---|---:|---:|')
    cands = [('table order', list(range(n))),
             ('nested dissection (first version)', nested_dissection(n, adj, False)),
             ('nested dissection, balanced tie-break', nested_dissection(n, adj, True)),
             ('minimum degree', greedy(n, adj, lambda nb, h, i: (len(nb[i]), i))),
             ('minimum fill', greedy(n, adj, lambda nb, h, i: (fill_of(nb, i), len(nb[i]), i))),
             ('shallowest subtree first, then degree', greedy(n, adj, lambda nb, h, i: (h[i], len(nb[i]), i)))]
    for label, order in cands:
        print('| %s | %d | %d |' % ((label,) + height_and_fill(n, adj, order)))
    print()


if __name__ == '__main__':
    # config 4: 500 keyframes, tracks of 6 consecutive keyframes -> supernodes of 5 poses couple with their neighbours
    n = 100
    study('config 4 (stereo BA, 499 variable poses)', n, [[j for j in (i - 1, i + 1) if 0 <= j < n] for i in range(n)])
    # config 2: SE(2) pose graph, 10 poses per supernode, odometry + 100 loop closures (i, i + 200)
    d = synthetic.se2_pose_graph(1000, 100, seed=0)
    n = 100
    adj = [set() for _ in range(n)]
    for i, j in list(zip(d['odo_i'], d['odo_j'])) + list(zip(d['loop_i'], d['loop_j'])):
        a, b = int(i) // 10, int(j) // 10
        if a != b:
            adj[a].add(b); adj[b].add(a)
    study('config 2 (SE(2) pose graph, 1 000 poses, 100 loop closures)', n, [sorted(a) for a in adj])
