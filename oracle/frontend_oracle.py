"""TEST INFRASTRUCTURE ONLY: restatement of the reference's candidate selection and closure voting
 on plain arrays, used to check the C++ mirrors in
include/cgm/slam_frontend.hpp. Nothing in the product imports this module.

Follows, function by function:
  dijkstra                     g2o HyperDijkstra::shortestPaths (third party, recalled: SURVEY C12)
  find_vertices_scan_matching  src/slam/vertices_finder.cpp:61-80 (+ :35-59)
  find_sets_of_vertices        src/slam/vertices_finder.cpp:82-99
  find_closest_vertex          src/slam/vertices_finder.cpp:101-113
  add_neighboring_vertices     src/slam/graph_slam.cpp:356-382
  check_covariance             src/slam/graph_slam.cpp:311-354 (marginals supplied by the caller)
  closure_check                src/slam/closure_checker.cpp:47-139
  ClosureWindow                src/slam/closure_buffer.cpp:50-110 + graph_slam.cpp:487-538
Parity status: the reference holds no tests or golden vectors for these functions and its g2o is
absent, so parity is UNPINNED; sets are iterated in ascending id (g2o: pointer order).
"""
import heapq
import math

import numpy as np

from oracle import pgo_oracle as po

MAX_GRAPH_DIST_SM = 2.0
MIN_GRAPH_DIST_LC = 5.0
MAX_EUC_DIST_LC = 50.0
INF = float("inf")


def adjacency(edges):
    adj = {}
    for a, b in edges:
        adj.setdefault(a, []).append(b)
        adj.setdefault(b, []).append(a)
    return adj


def dist(poses, a, b):
    d = poses[a][:2] - poses[b][:2]
    return math.sqrt(d[0] * d[0] + d[1] * d[1])


def dijkstra(adj, source, cost, max_distance=INF, conditioner=1e-3):
    """Vertices reached from `source`: relax when d_new + conditioner < d_old and d_new < max."""
    best = {source: 0.0}
    visited = {source}
    heap = [(0.0, source)]
    while heap:
        d, u = heapq.heappop(heap)
        if d > best[u]:
            continue
        for z in adj.get(u, ()):
            c = cost(u, z)
            if c == INF:
                continue
            nd = d + c
            if nd + conditioner < best.get(z, INF) and nd < max_distance:
                best[z] = nd
                visited.add(z)
                heapq.heappush(heap, (nd, z))
    return visited


def find_vertices_scan_matching(poses, edges, cur):
    """poses: {id: (x, y, th)}. Returns the sorted candidate ids."""
    adj = adjacency(edges)
    cost = lambda a, b: dist(poses, a, b)
    near = dijkstra(adj, cur, cost, MAX_GRAPH_DIST_SM)
    not_far = dijkstra(adj, cur, cost, MIN_GRAPH_DIST_LC)
    lc = {v for v in poses if v not in not_far and dist(poses, cur, v) <= MAX_EUC_DIST_LC}
    return sorted((near | lc) - {cur})


def find_sets_of_vertices(edges, vset):
    adj = adjacency(edges)
    rest = set(vset)
    groups = []
    while rest:
        inside = set(rest)
        start = min(rest)
        group = dijkstra(adj, start, lambda a, b: 1.0 if a in inside and b in inside else INF)
        groups.append(sorted(group))
        rest -= group
    return sorted(groups)     # std::set<VertexSet>: lexicographic by id


def find_closest_vertex(poses, group, cur):
    best, arg = INF, None
    for v in sorted(group):
        d = dist(poses, cur, v)
        if d < best:
            best, arg = d, v
    return arg


def add_neighboring_vertices(poses, vset, cur, gap):
    out = set(vset)
    for vertex in sorted(vset):
        for direction in (1, -1):
            for i in range(1, gap + 1):
                v = vertex + direction * i
                if v in poses and v != cur:
                    if v in out:
                        break
                    out.add(v)
    return sorted(out)


def check_covariance(poses, vset, cur, marginals):
    """marginals: {id: 3x3 covariance with `cur` as the gauge}. Returns the surviving ids."""
    keep = []
    for v in sorted(vset):
        p = marginals[v]
        delta = po.se2_mul(po.se2_inv(np.asarray(poses[v])), np.asarray(poses[cur]))[0]
        h = []
        for k in (0, 1):
            x = delta[k]
            h.append(x - 1.0 if x - 1.0 > 0 else (x + 1.0 if x + 1.0 < 0 else 0.0))
        a, b, c, d = p[0, 0], p[0, 1], p[1, 0], p[1, 1]
        det = a * d - b * c
        d2 = (h[0] * (d * h[0] - b * h[1]) + h[1] * (-c * h[0] + a * h[1])) / det
        if not d2 > 5.99:
            keep.append(v)
    return keep


def _edge_chi2(pi, pj, z, info):
    e = po.se2_mul(po.se2_inv(np.asarray(z)), po.se2_mul(po.se2_inv(np.asarray(pi)), np.asarray(pj)))[0]
    return float(e @ info @ e)


def closure_check(poses, window, candidates, threshold, info):
    """candidates: [(from, to, (dx, dy, dth))]; window: ids of the floating vertices.
    Returns (inliers, chi2, per-candidate chi2 of the best hypothesis)."""
    best_inl, best_chi2 = 0, INF
    best = [INF] * len(candidates)
    win = sorted(window)
    for (a, b, z) in candidates:
        root = None
        if a in window:
            root = a
        if b in window:
            root = b
        if root == a:
            new_root = po.se2_mul(np.asarray(poses[b]), po.se2_inv(np.asarray(z)))[0]
        else:
            new_root = po.se2_mul(np.asarray(poses[a]), np.asarray(z))[0]
        motion = po.se2_mul(new_root, po.se2_inv(np.asarray(poses[root])))[0]
        moved = dict(poses)
        for v in win:
            moved[v] = po.se2_mul(motion, np.asarray(poses[v]))[0]
        temp = [_edge_chi2(moved[i], moved[j], zz, info) for (i, j, zz) in candidates]
        inl = sum(1 for c in temp if c < threshold)
        tot = 0.0
        for c in temp:
            if c < threshold:
                tot += c
        if inl > best_inl or (inl == best_inl and tot < best_chi2):
            best_inl, best_chi2, best = inl, tot, temp
    return best_inl, best_chi2, best


class ClosureWindow:
    """ClosureBuffer driven as GraphSLAM drives it: add (edges, vertex); checkList; updateList."""

    def __init__(self, window):
        self.window = window
        self.vertices = []      # [id, time] in insertion order
        self.edges = []         # (serial, vertex id)
        self.serial = 0

    def step(self, vertex, n_edges):
        if n_edges > 0:
            for _ in range(n_edges):
                self.edges.append((self.serial, vertex))
                self.serial += 1
            self.vertices.append([vertex, 0])
        vote = any(t == self.window - 1 for _, t in self.vertices)
        for vt in self.vertices:
            vt[1] += 1
        for v, t in list(self.vertices):
            if t >= self.window:
                # removeVertex: the map entry decides whether anything is removed
                self.edges = [(s, ev) for (s, ev) in self.edges if ev != v]
                self.vertices = [vt for vt in self.vertices if vt[0] != v]
        ids = {v for v, _ in self.vertices}
        return vote, len(self.edges), len(ids), [tuple(vt) for vt in self.vertices]


class MRClosureWindows:
    """MRClosureBuffer (mr_closure_buffer.cpp:30-118): one ClosureBuffer per peer robot. State per
    peer: vertex list [[id, age]] in insertion order (a re-inserted vertex is listed again with
    age 0, closure_buffer.cpp addVertex), the id map, and the candidate edges (serial, vertex)."""

    def __init__(self):
        self.peers = {}
        self.serial = 0
        self.edges_of = {}

    def _remove_vertex(self, st, v):
        if v not in st["ids"]:
            return
        st["ids"].discard(v)
        st["edges"] = [(s, ev) for (s, ev) in st["edges"] if ev != v]
        st["list"] = [vt for vt in st["list"] if vt[0] != v]

    def insert(self, robot, vertex, n_edges):
        new = []
        for _ in range(n_edges):
            new.append((self.serial, vertex))
            self.serial += 1
        self.edges_of.setdefault(vertex, []).extend(new)
        st = self.peers.setdefault(robot, {"list": [], "ids": set(), "edges": []})
        st["list"].append([vertex, 0])
        st["ids"].add(vertex)
        st["edges"].extend(new)

    def remove(self, robot, vertex):
        st = self.peers.get(robot)
        if st is None:
            return
        self._remove_vertex(st, vertex)
        gone = set(self.edges_of.get(vertex, []))
        st["edges"] = [e for e in st["edges"] if e not in gone]
        if not st["ids"]:
            del self.peers[robot]

    def update(self, window):
        for robot in sorted(self.peers):
            st = self.peers[robot]
            for vt in st["list"]:
                vt[1] += 1
            for v, t in list(st["list"]):
                if t >= window:
                    self._remove_vertex(st, v)
            if not st["ids"]:
                del self.peers[robot]

    def state(self):
        return [(r, len(set(st["edges"])), len(st["ids"]), [tuple(vt) for vt in st["list"]])
                for r, st in sorted(self.peers.items())]

