// Host mirrors of the callers on either side of the two hot paths:
// the classes that decide WHICH vertices get matched and WHICH loop closures are accepted, with
// the reference's names, members and argument meaning, written against the g2o compatibility
// layer (g2o_compat.hpp) so that they drive the GPU solver / matcher through the same calls the
// reference makes:
//
//   GraphManipulator, CovarianceEstimator   src/slam/graph_manipulator.{h,cpp}:62-155
//   VerticesFinder (+ cost functors)        src/slam/vertices_finder.{h,cpp}:35-112
//   ClosureBuffer                           src/slam/closure_buffer.{h,cpp}
//   LoopClosureChecker                      src/slam/closure_checker.{h,cpp}:33-139
//   checkCovariance / addNeighboringVertices (GraphSLAM members, src/slam/graph_slam.cpp:311-382)
//
// All of it is sequential host logic on sets of a few tens of vertices / edges (the reference runs
// it once per keyframe under graphMutex); the arithmetic it triggers -- optimize(1), marginal
// covariance blocks, EdgeSE2 errors -- is the solver's. Sets iterate in ascending id / insertion
// order (g2o: pointer order), which fixes the order-dependent results.
#ifndef CGM_SLAM_FRONTEND_HPP
#define CGM_SLAM_FRONTEND_HPP

#include <cassert>
#include <limits>
#include <list>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../g2o_compat/g2o_compat.hpp"

// ---- graph_manipulator ------------------------------------------------------------------------------
struct GraphManipulator {
  explicit GraphManipulator(g2o::SparseOptimizer* optimizer) : _statePushed(false), _optimizer(optimizer) {}
  void setVertices(g2o::OptimizableGraph::VertexSet& vertices) {
    if (_statePushed) popState();
    _vertices = vertices;
  }
  void setGauge(g2o::OptimizableGraph::VertexSet& gauge) { _gauge = gauge; }
  void setGauge(g2o::OptimizableGraph::Vertex* gauge) {
    _gauge.clear();
    _gauge.insert(gauge);
  }
  void setEdges(g2o::OptimizableGraph::EdgeSet& edges) { _edges = edges; }

 protected:
  static g2o::OptimizableGraph::Vertex* ov(g2o::HyperGraph::Vertex* v) {
    return static_cast<g2o::OptimizableGraph::Vertex*>(v);
  }
  void pushState() {  // every vertex of the graph: estimate stack + fixed flag
    assert(!_statePushed && "state already pushed");
    _fixes.clear();
    for (auto& kv : _optimizer->vertices()) {
      ov(kv.second)->push();
      _fixes[kv.second] = ov(kv.second)->fixed();
    }
    _statePushed = true;
  }
  void popState() {
    assert(_statePushed && "no vertices to be restored");
    for (auto& kv : _optimizer->vertices()) {
      ov(kv.second)->pop();
      ov(kv.second)->setFixed(_fixes[kv.second]);
    }
    _statePushed = false;
  }
  void fixGauge() {  // gauge fixed, EVERY other vertex free (also the robot's first pose)
    assert(_statePushed && "trying to fix non-saved vertices");
    for (auto& kv : _optimizer->vertices()) ov(kv.second)->setFixed(_gauge.count(kv.second) != 0);
  }
  void optimize() {
    if (_edges.size()) _optimizer->initializeOptimization(_edges);
    else _optimizer->initializeOptimization();
    _optimizer->computeInitialGuess();
    _optimizer->optimize(1);
  }
  g2o::OptimizableGraph::VertexSet _vertices, _gauge;
  g2o::OptimizableGraph::EdgeSet _edges;
  std::map<g2o::HyperGraph::Vertex*, bool> _fixes;
  bool _statePushed;
  g2o::SparseOptimizer* _optimizer;
};

struct CovarianceEstimator : public GraphManipulator {
  explicit CovarianceEstimator(g2o::SparseOptimizer* optimizer) : GraphManipulator(optimizer) {}
  void compute() {
    pushState();
    fixGauge();
    optimize();
    std::vector<std::pair<int, int> > blockIndices;
    for (g2o::HyperGraph::Vertex* v : _vertices)
      if (ov(v)->hessianIndex() >= 0)
        blockIndices.push_back(std::make_pair(ov(v)->hessianIndex(), ov(v)->hessianIndex()));
    if (blockIndices.size()) _optimizer->computeMarginals(spinv, blockIndices);
    popState();
  }
  // valid until the next initializeOptimization re-assigns the hessian indices
  Eigen::MatrixXd getCovariance(g2o::OptimizableGraph::Vertex* v) {
    assert(!_gauge.count(v) && "trying to get covariance from fixed vertex");
    assert(_vertices.count(v) && "trying to get covariance from a vertex not contained in vertices");
    Eigen::MatrixXd cov;
    if (v->hessianIndex() >= 0) {
      Eigen::MatrixXd* b = spinv.block(v->hessianIndex(), v->hessianIndex());
      if (b) cov = *b;
    }
    return cov;
  }

 protected:
  g2o::SparseBlockMatrix<Eigen::MatrixXd> spinv;
};

// ---- vertices_finder --------------------------------------------------------------------------------
inline double distanceSE2(const g2o::SE2& r1, const g2o::SE2& r2) {
  const Eigen::Vector2d t = r1.translation() - r2.translation();
  return std::sqrt(t.x() * t.x() + t.y() * t.y());
}

inline double vertexDistance(g2o::HyperGraph::Vertex* v1, g2o::HyperGraph::Vertex* v2) {
  g2o::VertexSE2* a = dynamic_cast<g2o::VertexSE2*>(v1);
  g2o::VertexSE2* b = dynamic_cast<g2o::VertexSE2*>(v2);
  return a && b ? distanceSE2(a->estimate(), b->estimate()) : std::numeric_limits<double>::max();
}

// cost of an EdgeSE2 = Euclidean distance between its vertices' current estimates
struct MyCostFunction : public g2o::HyperDijkstra::CostFunction {
  double operator()(g2o::HyperGraph::Edge* e, g2o::HyperGraph::Vertex* from, g2o::HyperGraph::Vertex* to) override {
    if (!dynamic_cast<g2o::EdgeSE2*>(e)) return std::numeric_limits<double>::max();
    return vertexDistance(from, to);
  }
};

// cost 1 inside a vertex set, impassable outside: connected components of the set
struct MyVSetCostFunction : public g2o::HyperDijkstra::CostFunction {
  explicit MyVSetCostFunction(const g2o::OptimizableGraph::VertexSet& vset) : _vset(vset) {}
  double operator()(g2o::HyperGraph::Edge* e, g2o::HyperGraph::Vertex* from, g2o::HyperGraph::Vertex* to) override {
    if (!dynamic_cast<g2o::EdgeSE2*>(e) || !_vset.count(from) || !_vset.count(to))
      return std::numeric_limits<double>::max();
    return 1.0;
  }
  g2o::OptimizableGraph::VertexSet _vset;
};

#define MAX_GRAPH_DIST_SM 2.0   // max distance in the graph to find neighbours of a vertex
#define MIN_GRAPH_DIST_LC 5.0   // min distance in the graph for a loop-closure candidate
#define MAX_EUC_DIST_LC 50.0    // max Euclidean distance for a loop-closure candidate

class VerticesFinder {
 public:
  explicit VerticesFinder(g2o::SparseOptimizer* graph) : _graph(graph) {}

  void findVerticesInDistance(g2o::OptimizableGraph::VertexSet& vset, g2o::HyperGraph::Vertex* currentVertex,
                              double graphdist) {
    g2o::HyperDijkstra hd(_graph);
    MyCostFunction mcf;
    hd.shortestPaths(currentVertex, &mcf, graphdist);
    vset = hd.visited();
  }
  // everything NOT within `graphdist` along the graph but within MAX_EUC_DIST_LC as the crow flies
  void findVerticesLoopClosing(g2o::OptimizableGraph::VertexSet& vset, g2o::HyperGraph::Vertex* currentVertex,
                               double graphdist) {
    g2o::HyperDijkstra hd(_graph);
    MyCostFunction mcf;
    hd.shortestPaths(currentVertex, &mcf, graphdist);
    const g2o::OptimizableGraph::VertexSet& near = hd.visited();
    for (auto& kv : _graph->vertices())
      if (!near.count(kv.second) && vertexDistance(currentVertex, kv.second) <= MAX_EUC_DIST_LC)
        vset.insert(kv.second);
  }
  void findVerticesScanMatching(g2o::HyperGraph::Vertex* currentVertex, g2o::OptimizableGraph::VertexSet& vset) {
    findVerticesInDistance(vset, currentVertex, MAX_GRAPH_DIST_SM);
    g2o::OptimizableGraph::VertexSet vsetlc;
    findVerticesLoopClosing(vsetlc, currentVertex, MIN_GRAPH_DIST_LC);
    vset.insert(vsetlc.begin(), vsetlc.end());
    vset.erase(currentVertex);
  }
  // groups of vertices of `mixedvset` connected through edges that stay inside the set
  void findSetsOfVertices(g2o::OptimizableGraph::VertexSet& mixedvset,
                          std::set<g2o::OptimizableGraph::VertexSet>& setOfVSet) {
    setOfVSet.clear();
    g2o::OptimizableGraph::VertexSet rest = mixedvset;
    while (!rest.empty()) {
      g2o::HyperDijkstra hd(_graph);
      MyVSetCostFunction mcf(rest);
      hd.shortestPaths(*rest.begin(), &mcf);
      const g2o::OptimizableGraph::VertexSet group = hd.visited();
      setOfVSet.insert(group);
      for (g2o::HyperGraph::Vertex* v : group) rest.erase(v);
    }
  }
  // first vertex (ascending id) at minimal Euclidean distance
  g2o::OptimizableGraph::Vertex* findClosestVertex(g2o::OptimizableGraph::VertexSet& vset,
                                                   g2o::HyperGraph::Vertex* currentVertex) {
    double best = std::numeric_limits<double>::max();
    g2o::HyperGraph::Vertex* closest = nullptr;
    for (g2o::HyperGraph::Vertex* v : vset) {
      const double d = vertexDistance(currentVertex, v);
      if (d < best) {
        best = d;
        closest = v;
      }
    }
    return static_cast<g2o::OptimizableGraph::Vertex*>(closest);
  }

 protected:
  g2o::SparseOptimizer* _graph;
};

// ---- the two candidate filters of GraphSLAM::findConstraints ------------------------------------------
namespace cgm {

// GraphSLAM::checkCovariance (graph_slam.cpp:311-354): keep a candidate iff the current pose lies
// within the 95 % ellipse (chi2 5.99, 2 dof) of its xy marginal, measured in the candidate's frame
// after shrinking the offset by a 1 m perception range per axis. The marginals are those of the
// whole graph re-optimised once with `lastVertex` as the gauge (CovarianceEstimator).
inline void checkCovariance(g2o::SparseOptimizer* graph, g2o::VertexSE2* lastVertex,
                            g2o::OptimizableGraph::VertexSet& vset) {
  CovarianceEstimator ce(graph);
  ce.setVertices(vset);
  ce.setGauge(lastVertex);
  ce.compute();
  const g2o::OptimizableGraph::VertexSet tmp = vset;
  for (g2o::HyperGraph::Vertex* hv : tmp) {
    g2o::VertexSE2* v = static_cast<g2o::VertexSE2*>(hv);
    const Eigen::MatrixXd P = ce.getCovariance(v);
    if (P.rows() < 2) continue;  // no marginal (fixed / inactive): the reference would assert
    const g2o::SE2 delta = v->estimate().inverse() * lastVertex->estimate();
    double hx = delta.translation().x(), hy = delta.translation().y();
    const double range = 1.0;
    hx = hx - range > 0 ? hx - range : (hx + range < 0 ? hx + range : 0.0);
    hy = hy - range > 0 ? hy - range : (hy + range < 0 ? hy + range : 0.0);
    // d2 = h^T Pxy^-1 h with the closed-form 2x2 inverse
    const double a = P(0, 0), b = P(0, 1), c = P(1, 0), d = P(1, 1), det = a * d - b * c;
    const double d2 = (hx * (d * hx - b * hy) + hy * (-c * hx + a * hy)) / det;
    if (d2 > 5.99) vset.erase(hv);
  }
}

// GraphSLAM::addNeighboringVertices (graph_slam.cpp:356-382): grow every candidate by up to `gap`
// ids in both directions, stopping at the first id already in the set; never the current vertex.
inline void addNeighboringVertices(g2o::SparseOptimizer* graph, g2o::VertexSE2* lastVertex,
                                   g2o::OptimizableGraph::VertexSet& vset, int gap) {
  const g2o::OptimizableGraph::VertexSet temp = vset;
  for (g2o::HyperGraph::Vertex* vertex : temp)
    for (int dir = 1; dir >= -1; dir -= 2)
      for (int i = 1; i <= gap; ++i) {
        g2o::OptimizableGraph::Vertex* v = graph->vertex(vertex->id() + dir * i);
        if (v && v->id() != lastVertex->id()) {
          if (vset.count(v)) break;
          vset.insert(v);
        }
      }
}

}  // namespace cgm

// ---- closure_buffer -----------------------------------------------------------------------------------
struct VertexTime {
  int time;
  g2o::OptimizableGraph::Vertex* v;
};

// Sliding window of candidate loop closures: the vertices that produced candidates, each with an
// age in keyframes, and the candidate edges attached to them.
struct ClosureBuffer {
  void addEdge(g2o::OptimizableGraph::Edge* e) { _eset.insert(e); }
  void removeEdge(g2o::OptimizableGraph::Edge* e) { _eset.erase(e); }
  void addEdgeSet(g2o::OptimizableGraph::EdgeSet& eset) {
    for (g2o::HyperGraph::Edge* e : eset) addEdge(static_cast<g2o::OptimizableGraph::Edge*>(e));
  }
  void removeEdgeSet(g2o::OptimizableGraph::EdgeSet& eset) {
    for (g2o::HyperGraph::Edge* e : eset) removeEdge(static_cast<g2o::OptimizableGraph::Edge*>(e));
  }
  void addVertex(g2o::OptimizableGraph::Vertex* v) {
    _vmap.insert(std::make_pair(v->id(), v));
    const VertexTime vt = {0, v};
    _vertexList.push_back(vt);
  }
  void removeVertex(g2o::OptimizableGraph::Vertex* v) {  // with every candidate edge touching it
    if (!_vmap.erase(v->id())) return;
    const g2o::OptimizableGraph::EdgeSet tmp = _eset;
    for (g2o::HyperGraph::Edge* e : tmp)
      for (size_t i = 0; i < e->vertices().size(); ++i)
        if (e->vertex(i)->id() == v->id()) _eset.erase(e);
    _vertexList.remove_if([v](const VertexTime& vt) { return vt.v->id() == v->id(); });
  }
  g2o::OptimizableGraph::Vertex* findVertex(int idVertex) {
    g2o::OptimizableGraph::VertexIDMap::iterator it = _vmap.find(idVertex);
    return it == _vmap.end() ? nullptr : static_cast<g2o::OptimizableGraph::Vertex*>(it->second);
  }
  g2o::OptimizableGraph::EdgeSet& edgeSet() { return _eset; }
  g2o::OptimizableGraph::VertexIDMap& vertices() { return _vmap; }
  std::list<VertexTime>& vertexList() { return _vertexList; }
  // one keyframe passes: age everything, drop what reached the window size
  void updateList(int windowSize) {
    for (VertexTime& vt : _vertexList) vt.time++;
    const std::list<VertexTime> tmp(_vertexList);
    for (const VertexTime& vt : tmp)
      if (vt.time >= windowSize) removeVertex(vt.v);
  }
  // true when some vertex is about to leave the window: time to vote
  bool checkList(int windowSize) {
    for (const VertexTime& vt : _vertexList)
      if (vt.time == windowSize - 1) return true;
    return false;
  }

 protected:
  std::list<VertexTime> _vertexList;
  g2o::OptimizableGraph::EdgeSet _eset;
  g2o::OptimizableGraph::VertexIDMap _vmap;
};

// ---- closure_checker ------------------------------------------------------------------------------------
// Consensus among candidate closures: for each candidate in turn, move the window's vertices
// rigidly so that the candidate has zero error, evaluate chi2 of ALL candidates, count those below
// the inlier threshold; keep the hypothesis with most inliers (ties: lowest summed chi2).
class LoopClosureChecker {
 public:
  typedef std::map<g2o::OptimizableGraph::Edge*, double, g2o::HyperGraph::Vertex::EdgeLess> EdgeDoubleMap;
  LoopClosureChecker() : _bestChi2(0.0), _bestInliers(0), _inlierThreshold(0.0) {}
  void init(g2o::OptimizableGraph::VertexIDMap& movableRegion, g2o::OptimizableGraph::EdgeSet& closingEdges,
            double inlierThreshold) {
    _localVertices = movableRegion;
    _closuresToCheck = closingEdges;
    _inlierThreshold = inlierThreshold;
    _bestInliers = 0;
    _bestChi2 = std::numeric_limits<double>::max();
    _tempResult.clear();
    _bestResult.clear();
    for (g2o::HyperGraph::Edge* he : closingEdges) {
      g2o::OptimizableGraph::Edge* e = static_cast<g2o::OptimizableGraph::Edge*>(he);
      _tempResult[e] = std::numeric_limits<double>::max();
      _bestResult[e] = std::numeric_limits<double>::max();
    }
  }
  int inliers() const { return _bestInliers; }
  EdgeDoubleMap& closures() { return _bestResult; }
  double chi2() const { return _bestChi2; }

  void check(const std::string& matchingType = "2dPose") {
    assert(matchingType == "2dPose" && "wrong matching strategy selected");
    (void)matchingType;
    for (g2o::HyperGraph::Edge* he : _closuresToCheck) {
      applyZeroErrorTransform(static_cast<g2o::EdgeSE2*>(he));
      int inliers = 0;
      double totalChi2 = 0.0;
      for (auto& kv : _tempResult)
        if (kv.second < _inlierThreshold) {
          ++inliers;
          totalChi2 += kv.second;
        }
      if (inliers > _bestInliers || (inliers == _bestInliers && totalChi2 < _bestChi2)) {
        _bestInliers = inliers;
        _bestChi2 = totalChi2;
        _bestResult = _tempResult;
      }
    }
  }

 protected:
  void applyZeroErrorTransform(g2o::EdgeSE2* hypothesis) {
    g2o::VertexSE2* vfrom = static_cast<g2o::VertexSE2*>(hypothesis->vertex(0));
    g2o::VertexSE2* vto = static_cast<g2o::VertexSE2*>(hypothesis->vertex(1));
    // the end of the edge that lies in the floating window (the `to` end wins if both do)
    g2o::VertexSE2* root = nullptr;
    if (_localVertices.count(vfrom->id())) root = vfrom;
    if (_localVertices.count(vto->id())) root = vto;
    assert(root && "the loop closure does not have any vertex in the floating part of the map");
    if (!root) return;
    const g2o::SE2 newRootPose = root == vfrom ? vto->estimate() * hypothesis->measurement().inverse()
                                               : vfrom->estimate() * hypothesis->measurement();
    const g2o::SE2 motion = newRootPose * root->estimate().inverse();
    for (auto& kv : _localVertices) {
      g2o::VertexSE2* v = static_cast<g2o::VertexSE2*>(kv.second);
      v->push();
      v->setEstimate(motion * v->estimate());
    }
    for (auto& kv : _tempResult) {
      kv.first->computeError();
      kv.second = kv.first->chi2();
    }
    for (auto& kv : _localVertices) static_cast<g2o::VertexSE2*>(kv.second)->pop();
  }

  EdgeDoubleMap _bestResult;
  double _bestChi2;
  int _bestInliers;
  EdgeDoubleMap _tempResult;
  g2o::OptimizableGraph::VertexIDMap _localVertices;
  g2o::OptimizableGraph::EdgeSet _closuresToCheck;
  double _inlierThreshold;
};

#endif
