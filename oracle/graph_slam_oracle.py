"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's single-robot keyframe pipeline,
GraphSLAM (src/slam/graph_slam.cpp: setInitialData :89-142, addDataSM :197-264, findConstraints
:388-485, addClosures / checkClosures / updateClosures :487-560, optimize :561-575) as driven by
src/srslam.cpp:190-221, assembled from the other oracles: the CPU scan matcher
(oracle/scan_matcher_oracle.py over oracle/bindings.py), the pose-graph oracle (pgo_oracle.py) and
the candidate-selection / closure-voting oracle (frontend_oracle.py). Used to check the GPU-backed
GraphSLAM mirror (include/cgm/graph_slam.hpp) on keyframes of the reference's own bag files.
Parity status: UNPINNED (the reference cannot run here: no ROS, no g2o; it ships no recorded
outputs); sets are iterated in ascending id. Nothing in the product imports this module."""
import numpy as np

from oracle import frontend_oracle as fo
from oracle import pgo_oracle as po
from oracle import scan_matcher_oracle as smo

ODOM_INFO = (100.0, 0.0, 0.0, 100.0, 0.0, 1000.0)
SM_INFO = (1000.0, 0.0, 0.0, 1000.0, 0.0, 10000.0)


class GraphSlamOracle:
    def __init__(self, lib, orc, geom, laser_pose, window=10, max_score=0.15, inlier_threshold=2.0,
                 min_inliers=7, base_id=10000, id_robot=0):
        self.lib, self.orc = lib, orc
        self.geom, self.laser_pose = geom, laser_pose
        self.window, self.max_score = window, max_score
        self.inlier_threshold, self.min_inliers = inlier_threshold, min_inliers
        self.base_id, self.id_robot = base_id, id_robot
        self.vertices = {}          # id -> dict(id, pose, ranges, ...)
        self.edges = []             # graph edges in insertion order: dict(frm, to, z, info)
        self.running_vertex = self.running_edge = 0
        self.last = None
        self.last_odom = None
        self.buf_vertices = []      # [id, time]
        self.candidates = []        # dict(serial, frm, to, z, in_graph)
        self.events = []

    # ---- helpers ------------------------------------------------------------------------------------
    def _vertex(self, vid, pose, ranges):
        first, step, max_range = self.geom
        self.vertices[vid] = dict(id=vid, pose=np.asarray(pose, dtype=np.float64), ranges=ranges,
                                  first_angle=first, step=step, max_range=max_range,
                                  laser_pose=self.laser_pose)
        return self.vertices[vid]

    def _arrays(self):
        ids = sorted(self.vertices)
        index = {v: k for k, v in enumerate(ids)}
        poses = np.array([self.vertices[v]["pose"] for v in ids])
        eij = np.array([[index[e["frm"]], index[e["to"]]] for e in self.edges], dtype=np.int32).reshape(-1, 2)
        meas = np.array([e["z"] for e in self.edges]).reshape(-1, 3)
        info = np.array([e["info"] for e in self.edges]).reshape(-1, 6)
        return ids, index, poses, eij, meas, info

    def _optimize(self, n):
        ids, index, poses, eij, meas, info = self._arrays()
        res = po.gauss_newton(poses, eij, meas, info, [index[self.id_robot * self.base_id]], n)
        for k, v in enumerate(ids):
            self.vertices[v]["pose"] = res.poses[k].copy()

    def _mine(self, vid):
        return vid // self.base_id == self.id_robot

    # ---- the pipeline -------------------------------------------------------------------------------
    def set_initial_data(self, odom, ranges):
        self.last_odom = np.asarray(odom, dtype=np.float64)
        self.last = self._vertex(self.id_robot * self.base_id, odom, ranges)

    def add_data_sm(self, odom, ranges):
        odom = np.asarray(odom, dtype=np.float64)
        disp = po.se2_mul(po.se2_inv(self.last_odom), odom)[0]
        est = po.se2_mul(self.last["pose"], disp)[0]
        self.running_vertex += 1
        v = self._vertex(self.running_vertex + self.id_robot * self.base_id, est, ranges)
        vset = [self.last]
        for j in range(1, 6):
            vj = self.vertices.get(self.last["id"] - j)
            if vj is None:
                break
            vset.append(vj)
        ok, t = smo.close_scan_matching(self.lib, self.orc, vset, self.last, v, self.max_score)
        self.running_edge += 1
        if ok:
            self.edges.append(dict(frm=self.last["id"], to=v["id"], z=np.array(t), info=SM_INFO))
            self.events.append(("S", self.last["id"], v["id"], np.array(t)))
        else:
            self.edges.append(dict(frm=self.last["id"], to=v["id"], z=disp, info=ODOM_INFO))
            self.events.append(("O", self.last["id"], v["id"], disp))
        self.last_odom = odom
        self.last = v

    def find_constraints(self):
        self._optimize(1)
        cur = self.last["id"]
        poses = {v: d["pose"] for v, d in self.vertices.items()}
        pairs = [(e["frm"], e["to"]) for e in self.edges]
        vset = fo.find_vertices_scan_matching(poses, pairs, cur)
        # checkCovariance: whole graph re-optimised once with the current vertex as the gauge
        if vset:
            ids, index, parr, eij, meas, info = self._arrays()
            guess = po.initial_guess(parr, eij, meas, [index[cur]])
            res = po.gauss_newton(guess, eij, meas, info, [index[cur]], 1)
            hidx = po.hessian_index(len(ids), [index[cur]])
            covs = po.marginals(res, hidx, [(index[v], index[v]) for v in vset])
            vset = fo.check_covariance(poses, vset, cur, {v: c for v, c in zip(vset, covs)})
        vset = fo.add_neighboring_vertices(poses, vset, cur, 8)
        new_candidates = []
        for group in fo.find_sets_of_vertices(pairs, vset):
            closest = fo.find_closest_vertex(poses, group, cur)
            if closest == cur - 1:
                continue
            gv = [self.vertices[v] for v in group]
            if not self._mine(closest) or abs(cur - closest) > 10:
                results = smo.scan_matching_lc(self.lib, self.orc, gv, self.vertices[closest], self.last,
                                               self.max_score)
                if len(results):
                    for r in results:
                        self.running_edge += 1
                        c = dict(serial=10**9 + self.running_edge, frm=closest, to=cur, z=np.array(r),
                                 in_graph=False)
                        new_candidates.append(c)
                        self.events.append(("L", closest, cur, np.array(r)))
                else:
                    self.events.append(("R", closest, cur, np.zeros(3)))
            else:
                ok, t = smo.close_scan_matching(self.lib, self.orc, gv, self.vertices[closest], self.last,
                                                self.max_score)
                if ok:
                    self.running_edge += 1
                    self.edges.append(dict(frm=closest, to=cur, z=np.array(t), info=SM_INFO))
                    self.events.append(("C", closest, cur, np.array(t)))
                else:
                    self.events.append(("R", closest, cur, np.zeros(3)))
        if new_candidates:
            self.candidates += new_candidates
            self.buf_vertices.append([cur, 0])
        self._check_closures()
        # updateClosures
        for vt in self.buf_vertices:
            vt[1] += 1
        for v, t in list(self.buf_vertices):
            if t >= self.window:
                self.candidates = [c for c in self.candidates if c["frm"] != v and c["to"] != v]
                self.buf_vertices = [vt for vt in self.buf_vertices if vt[0] != v]

    def _check_closures(self):
        if not any(t == self.window - 1 for _, t in self.buf_vertices):
            return
        poses = {v: d["pose"] for v, d in self.vertices.items()}
        cands = sorted(self.candidates, key=lambda c: c["serial"])
        window = {v for v, _ in self.buf_vertices}
        info = np.diag([1000.0, 1000.0, 10000.0])
        inl, chi2, per = fo.closure_check(poses, window, [(c["frm"], c["to"], c["z"]) for c in cands],
                                          self.inlier_threshold, info)
        if inl < self.min_inliers:
            return
        for c, x in zip(cands, per):
            if x < self.inlier_threshold and not c["in_graph"]:
                c["in_graph"] = True
                c["serial"] = len(self.edges)          # the graph's insertion serial
                self.edges.append(dict(frm=c["frm"], to=c["to"], z=c["z"], info=SM_INFO))
                self.events.append(("A", c["frm"], c["to"], c["z"]))

    def optimize(self, n):
        self._optimize(n)

    def take_events(self):
        ev, self.events = self.events, []
        return ev
