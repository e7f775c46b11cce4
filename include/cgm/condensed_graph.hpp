// Condensed graphs: host mirrors of
//   CondensedGraphCreator   src/mrslam/condensed_graph/condensed_graph_creator.{h,cpp}:33-66
//   CondensedGraphBuffer    src/mrslam/condensed_graph/condensed_graph_buffer.{h,cpp}:94-510
//  with the reference's names, members and argument meaning, written against
// the g2o compatibility layer. What they trigger on the GPU: GraphManipulator::optimize() -- one
// Gauss-Newton iteration of the robot's own edges with only the gauge fixed -- and
// EdgeLabeler::labelEdges() -- marginal blocks of the separators + the unscented transform of
// every star edge (pgo_label_star_edges).
//
// A robot keeps, per peer robot: the vertices of the peer that closed loops with it ("in"
// closures), its own vertices the peer asked about ("out" closures), the star it computed for the
// peer (level peer + 1 edges in its own graph, excluded from every optimisation of its own data)
// and the star it received from the peer (level 0 edges: they take part in optimize()).
// Sets iterate in ascending id / insertion order (g2o: pointer order), as in slam_frontend.hpp.
#ifndef CGM_CONDENSED_GRAPH_HPP
#define CGM_CONDENSED_GRAPH_HPP

#include <cassert>
#include <iostream>
#include <limits>
#include <map>
#include <set>
#include <vector>

#include "slam_frontend.hpp"

// Star of measurements gauge -> v over a vertex set, from the edges given with setEdges().
struct CondensedGraphCreator : public GraphManipulator {
  explicit CondensedGraphCreator(g2o::SparseOptimizer* optimizer) : GraphManipulator(optimizer) {}

  // Computes a condensed graph linking the gauge with the rest of the vertices. The new edges are
  // owned by the caller (the buffer hands them to the graph).
  void compute() {
    assert(_gauge.size() == 1 && "cannot create a condensed graph with multiple gauges");
    g2o::OptimizableGraph::Vertex* gauge = ov(*_gauge.begin());
    assert(_vertices.count(gauge) && "gauge must be contained in vertices");
    pushState();
    fixGauge();
    optimize();
    std::set<g2o::OptimizableGraph::Edge*> edgesToLabel;
    long long serial = next_serial();
    for (g2o::HyperGraph::Vertex* hv : _vertices) {
      if (hv->id() == gauge->id()) continue;
      g2o::EdgeSE2* e = new g2o::EdgeSE2;
      e->vertices()[0] = gauge;
      e->vertices()[1] = hv;
      e->setSerial(serial++);   // creation order = iteration order of the sets holding it
      _condensedGraph.insert(e);
      edgesToLabel.insert(e);
    }
    g2o::EdgeLabeler labeler(_optimizer);
    labeler.labelEdges(edgesToLabel);
    popState();
  }
  g2o::OptimizableGraph::EdgeSet& getCondensedGraph() { return _condensedGraph; }

 protected:
  // serials above everything the graph handed out so far
  long long next_serial() const {
    long long s = 1LL << 40;
    for (g2o::HyperGraph::Edge* e : _optimizer->edges()) s = e->serial() >= s ? e->serial() + 1 : s;
    return s;
  }
  g2o::OptimizableGraph::EdgeSet _condensedGraph;
};

struct CondensedGraphBuffer {
  typedef g2o::OptimizableGraph::VertexIDMap VertexIDMap;
  typedef g2o::OptimizableGraph::EdgeSet EdgeSet;

  explicit CondensedGraphBuffer(g2o::SparseOptimizer* optimizer) : _optimizer(optimizer) {}
  ~CondensedGraphBuffer() {
    for (auto& kv : _inClosures) delete kv.second;
    for (auto& kv : _outClosures) delete kv.second;
  }
  CondensedGraphBuffer(const CondensedGraphBuffer&) = delete;
  CondensedGraphBuffer& operator=(const CondensedGraphBuffer&) = delete;

  // vertices are added once; a vertex already listed keeps its entry (condensed_graph_buffer.cpp:129-171)
  void insertInClosure(int robot, VertexIDMap& vidmap) { merge(_inClosures, robot, vidmap); }
  void insertOutClosure(int robot, VertexIDMap& vidmap) { merge(_outClosures, robot, vidmap); }

  // the graph's edges minus the stars received from and built for other robots
  EdgeSet getMyEdges() {
    EdgeSet mine = _optimizer->edges();
    for (auto& kv : _inCondensedGraphs)
      for (g2o::HyperGraph::Edge* e : kv.second) mine.erase(e);
    for (auto& kv : _outCondensedGraphs)
      for (g2o::HyperGraph::Edge* e : kv.second) mine.erase(e);
    return mine;
  }

  // Gauge selection: argmin of a cost over the candidate vertices, ascending id, the first of equal
  // candidates wins (strict < in condensed_graph_buffer.cpp:270,305,337).
  //   optimal:  sum over the star it would yield of det(Omega_e^-1)   (:252-288)
  //   plain:    summed distance to the other candidates               (:290-316)
  //   centroid: distance to the centroid of the candidates            (:318-345)
  g2o::OptimizableGraph::Vertex* selectOptimalGauge(VertexIDMap* vertices) {
    EdgeSet own = getMyEdges();
    g2o::OptimizableGraph::VertexSet all;
    for (auto& kv : *vertices) all.insert(kv.second);
    return cheapest(vertices, [&](g2o::VertexSE2* candidate) {
      CondensedGraphCreator creator(_optimizer);
      creator.setVertices(all);
      creator.setGauge(candidate);
      creator.setEdges(own);
      creator.compute();
      EdgeSet star = creator.getCondensedGraph();
      const double cost = computeOverallUncertainty(star);
      for (g2o::HyperGraph::Edge* e : star) delete e;   // (the reference leaks them)
      return cost;
    });
  }
  g2o::OptimizableGraph::Vertex* selectGauge(VertexIDMap* vertices) {
    return cheapest(vertices, [&](g2o::VertexSE2* candidate) {
      double sum = 0;
      for (auto& kv : *vertices) {
        g2o::VertexSE2* other = static_cast<g2o::VertexSE2*>(kv.second);
        if (other->id() != candidate->id())
          sum += (candidate->estimate().inverse() * other->estimate()).translation().norm();
      }
      return sum;
    });
  }
  g2o::OptimizableGraph::Vertex* selectGaugeCentroid(VertexIDMap* vertices) {
    Eigen::Vector2d sum(.0, .0);
    for (auto& kv : *vertices) sum += static_cast<g2o::VertexSE2*>(kv.second)->estimate().translation();
    const Eigen::Vector2d centroid(sum.x() / vertices->size(), sum.y() / vertices->size());
    return cheapest(vertices, [&](g2o::VertexSE2* candidate) {
      const Eigen::Vector2d d = candidate->estimate().translation() - centroid;
      return d.norm();
    });
  }

  // The star robot `robot` gets: over the vertices it asked about (insertOutClosure), from this
  // robot's own edges only. It replaces the previous star for that robot in the graph (level
  // robot + 1). Nothing happens when the robot has not asked for any vertex yet.
  void computeCondensedGraph(int robot, bool optimal = false) {
    VertexIDMap* outvidmap = outClosures(robot);
    if (!outvidmap) return;
    // the previous star leaves the graph (and is deleted with it) before "my edges" are listed
    const int level = robot + 1;
    _outCondensedGraphs.erase(robot);
    removeSubgraph(level);
    EdgeSet myOwnEdges = getMyEdges();

    CondensedGraphCreator gc(_optimizer);
    g2o::OptimizableGraph::VertexSet subGraph;
    for (auto& kv : *outvidmap) subGraph.insert(kv.second);
    gc.setVertices(subGraph);
    g2o::OptimizableGraph::Vertex* gauge = optimal ? selectOptimalGauge(outvidmap) : selectGaugeCentroid(outvidmap);
    gc.setGauge(gauge);
    gc.setEdges(myOwnEdges);
    gc.compute();

    // addEdge gives every edge its place in the graph's insertion order (the key of an EdgeSet):
    // hand them over in creation order first, build the stored set afterwards
    std::vector<g2o::HyperGraph::Edge*> labeled(gc.getCondensedGraph().begin(), gc.getCondensedGraph().end());
    gc.getCondensedGraph().clear();
    for (g2o::HyperGraph::Edge* he : labeled) {
      g2o::OptimizableGraph::Edge* edge = static_cast<g2o::OptimizableGraph::Edge*>(he);
      edge->setLevel(level);
      _optimizer->addEdge(edge);
    }
    _outCondensedGraphs.insert(std::make_pair(robot, EdgeSet(labeled.begin(), labeled.end())));
  }

  // The star received from `robot` replaces the previous one (level 0: it takes part in optimize()).
  void insertEdgesFromRobot(int robot, EdgeSet& eset) {
    std::map<int, EdgeSet>::iterator itr = _inCondensedGraphs.find(robot);
    if (itr != _inCondensedGraphs.end()) {
      EdgeSet oldEdges = itr->second;
      for (g2o::HyperGraph::Edge* e : oldEdges) _optimizer->removeEdge(e);
      _inCondensedGraphs.erase(itr);
    }
    // (see computeCondensedGraph for the order of addEdge and set construction; `eset` itself is
    // emptied: its keys change when the graph adopts the edges)
    std::vector<g2o::HyperGraph::Edge*> fresh(eset.begin(), eset.end());
    eset.clear();
    for (g2o::HyperGraph::Edge* e : fresh) _optimizer->addEdge(static_cast<g2o::OptimizableGraph::Edge*>(e));
    _inCondensedGraphs.insert(std::make_pair(robot, EdgeSet(fresh.begin(), fresh.end())));
  }

  std::map<int, VertexIDMap*>& inClosures() { return _inClosures; }
  std::map<int, VertexIDMap*>& outClosures() { return _outClosures; }
  std::map<int, EdgeSet>& outCondensedGraphs() { return _outCondensedGraphs; }
  std::map<int, EdgeSet>& inCondensedGraphs() { return _inCondensedGraphs; }
  VertexIDMap* inClosures(int robot) { return find(_inClosures, robot); }
  VertexIDMap* outClosures(int robot) { return find(_outClosures, robot); }
  EdgeSet outCondensedGraph(int robot) { return find(_outCondensedGraphs, robot); }
  EdgeSet inCondensedGraph(int robot) { return find(_inCondensedGraphs, robot); }

  // sum over the edges of det(Omega^-1) (condensed_graph_buffer.cpp:172-180)
  static double computeOverallUncertainty(EdgeSet& edges) {
    double totalUncertainty = 0;
    for (g2o::HyperGraph::Edge* he : edges)
      totalUncertainty += static_cast<g2o::EdgeSE2*>(he)->information().inverse().determinant();
    return totalUncertainty;
  }

 protected:
  static g2o::OptimizableGraph::Vertex* ovx(g2o::HyperGraph::Vertex* v) {
    return static_cast<g2o::OptimizableGraph::Vertex*>(v);
  }
  template <class Cost>
  static g2o::OptimizableGraph::Vertex* cheapest(VertexIDMap* vertices, Cost cost) {
    g2o::OptimizableGraph::Vertex* best = 0;
    double best_cost = std::numeric_limits<double>::max();
    for (auto& kv : *vertices) {
      g2o::VertexSE2* candidate = static_cast<g2o::VertexSE2*>(kv.second);
      const double c = cost(candidate);
      if (c < best_cost) {
        best_cost = c;
        best = candidate;
      }
    }
    return best;
  }
  static void merge(std::map<int, VertexIDMap*>& into, int robot, VertexIDMap& vidmap) {
    std::map<int, VertexIDMap*>::iterator it = into.find(robot);
    if (it == into.end()) {
      into.insert(std::make_pair(robot, new VertexIDMap(vidmap)));
      return;
    }
    for (auto& kv : vidmap) it->second->insert(kv);   // std::map::insert keeps an existing entry
  }
  static VertexIDMap* find(std::map<int, VertexIDMap*>& in, int robot) {
    std::map<int, VertexIDMap*>::iterator it = in.find(robot);
    return it == in.end() ? 0 : it->second;
  }
  static EdgeSet find(std::map<int, EdgeSet>& in, int robot) {
    std::map<int, EdgeSet>::iterator it = in.find(robot);
    return it == in.end() ? EdgeSet() : it->second;
  }
  // every edge of that level leaves the graph (removeEdge deletes it); level 0 is never touched
  void removeSubgraph(int level) {
    if (level <= 0) return;
    EdgeSet tmp(_optimizer->edges());
    for (g2o::HyperGraph::Edge* he : tmp)
      if (static_cast<g2o::OptimizableGraph::Edge*>(he)->level() == level) _optimizer->removeEdge(he);
  }

  std::map<int, VertexIDMap*> _inClosures;
  std::map<int, VertexIDMap*> _outClosures;
  std::map<int, EdgeSet> _outCondensedGraphs;
  std::map<int, EdgeSet> _inCondensedGraphs;
  g2o::SparseOptimizer* _optimizer;
};

#endif
