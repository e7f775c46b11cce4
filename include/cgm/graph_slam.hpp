// Host mirror of the reference's single-robot pipeline object, GraphSLAM
// (src/slam/graph_slam.{h,cpp}: init :47-77, setInitialData :89-142, addData :144-195, addDataSM
// :197-264, findConstraints :388-485, addClosures / checkClosures / updateClosures :487-560,
// optimize :561-575), with the reference's member names and argument meaning, written against the
// mirrors of the classes it is made of: ScanMatcher (scan_matcher.hpp -> GPU matcher), the g2o
// compatibility layer (SparseOptimizer -> GPU solver), VerticesFinder / ClosureBuffer /
// LoopClosureChecker / CovarianceEstimator (slam_frontend.hpp). It exists so that the keyframe loop
// of src/srslam.cpp:190-221 can be replayed offline (tests/cpp/srslam_replay.cpp) and checked
// against the CPU oracle pipeline: which closures are found, between which vertices, and the
// resulting poses. Control logic only -- no arithmetic of its own. Not thread-safe (the reference
// guards every member with one mutex; a replay is single-threaded).
#ifndef CGM_GRAPH_SLAM_HPP
#define CGM_GRAPH_SLAM_HPP

#include <cstdlib>
#include <iostream>
#include <memory>
#include <vector>

#include "scan_matcher.hpp"
#include "slam_frontend.hpp"

class GraphSLAM {
 public:
  // one record per decision, for parity checks: kind 'O' odometry edge kept, 'S' scan-matched
  // odometry edge, 'C' close edge added, 'L' loop-closure candidate, 'A' closure accepted into the
  // graph, 'R' matching rejected
  struct Event {
    char kind;
    int from, to;
    g2o::SE2 measurement;
  };

  GraphSLAM()
      : _graph(new g2o::SparseOptimizer()), _idRobot(0), _baseId(10000), _runningVertexId(0),
        _runningEdgeId(0), _firstRobotPose(nullptr), _lastVertex(nullptr), _vf(_graph),
        windowLoopClosure(10), maxScore(0.15), inlierThreshold(2.0), minInliers(7) {}
  ~GraphSLAM() {
    // candidate edges that never entered the graph are ours
    for (g2o::EdgeSE2* e : _candidates)
      if (!_graph->edges().count(e)) delete e;
    delete _graph;
  }
  GraphSLAM(const GraphSLAM&) = delete;
  GraphSLAM& operator=(const GraphSLAM&) = delete;

  void init(double resolution, double kernelRadius, int windowLoopClosure_, double maxScore_,
            double inlierThreshold_, int minInliers_) {
    typedef g2o::BlockSolver<g2o::BlockSolverTraits<-1, -1> > SlamBlockSolver;
    typedef g2o::LinearSolverCSparse<SlamBlockSolver::PoseMatrixType> SlamLinearSolver;
    auto linearSolver = std::unique_ptr<SlamLinearSolver>(new SlamLinearSolver());
    linearSolver->setBlockOrdering(false);
    auto blockSolver = std::unique_ptr<SlamBlockSolver>(new SlamBlockSolver(std::move(linearSolver)));
    _graph->setAlgorithm(new g2o::OptimizationAlgorithmGaussNewton(std::move(blockSolver)));
    _graph->setVerbose(false);
    _closeMatcher.initializeKernel(resolution, kernelRadius);
    _closeMatcher.initializeGrid(Eigen::Vector2f(-15, -15), Eigen::Vector2f(15, 15), resolution);
    _LCMatcher.initializeKernel(0.1, 0.5);
    _LCMatcher.initializeGrid(Eigen::Vector2f(-35, -35), Eigen::Vector2f(35, 35), 0.1);
    windowLoopClosure = windowLoopClosure_;
    maxScore = maxScore_;
    inlierThreshold = inlierThreshold_;
    minInliers = minInliers_;
    _odominf = Eigen::Matrix3d::Identity();
    _odominf(0, 0) = _odominf(1, 1) = 100.0;
    _odominf(2, 2) = 1000.0;
    _SMinf = Eigen::Matrix3d::Identity();
    _SMinf(0, 0) = _SMinf(1, 1) = 1000.0;
    _SMinf(2, 2) = 10000.0;
  }

  void setIdRobot(int idRobot) { _idRobot = idRobot; }
  int idRobot() const { return _idRobot; }
  void setBaseId(int baseId) { _baseId = baseId; }
  int baseId() const { return _baseId; }
  bool isMyVertex(g2o::OptimizableGraph::Vertex* v) { return v->id() / baseId() == idRobot(); }

  void setInitialData(g2o::SE2 initialTruePose, g2o::SE2 initialOdom, g2o::RobotLaser* laser) {
    _lastOdom = initialOdom;
    firstVertex(initialTruePose, laser);
  }
  void setInitialData(g2o::SE2 initialOdom, g2o::RobotLaser* laser) {
    _lastOdom = initialOdom;
    firstVertex(initialOdom, laser);
  }

  // new keyframe + plain odometry edge
  void addData(g2o::SE2 currentOdom, g2o::RobotLaser* laser) {
    const g2o::SE2 displacement = _lastOdom.inverse() * currentOdom;
    g2o::VertexSE2* v = newVertex(displacement, laser);
    g2o::EdgeSE2* e = newEdge(_lastVertex, v, displacement, _odominf);
    _graph->addEdge(e);
    log('O', e);
    _lastOdom = currentOdom;
    _lastVertex = v;
  }

  // new keyframe + odometry edge refined by matching against the last (up to) six keyframes
  void addDataSM(g2o::SE2 currentOdom, g2o::RobotLaser* laser) {
    const g2o::SE2 displacement = _lastOdom.inverse() * currentOdom;
    g2o::VertexSE2* v = newVertex(displacement, laser);
    g2o::OptimizableGraph::VertexSet vset;
    vset.insert(_lastVertex);
    for (int j = 1; j <= 5; ++j) {
      g2o::OptimizableGraph::Vertex* vj = _graph->vertex(_lastVertex->id() - j);
      if (!vj) break;
      vset.insert(vj);
    }
    g2o::SE2 transf;
    const bool matched = _closeMatcher.closeScanMatching(vset, _lastVertex, v, &transf, maxScore);
    g2o::EdgeSE2* e = matched ? newEdge(_lastVertex, v, transf, _SMinf)
                              : newEdge(_lastVertex, v, displacement, _odominf);
    _graph->addEdge(e);
    log(matched ? 'S' : 'O', e);
    _lastOdom = currentOdom;
    _lastVertex = v;
  }

  g2o::SparseOptimizer* graph() { return _graph; }
  g2o::VertexSE2* lastVertex() { return _lastVertex; }
  g2o::SE2 lastOdom() { return _lastOdom; }

  void findConstraints() {
    // the graph is optimised once first so that the last added edge is satisfied
    _graph->initializeOptimization();
    _graph->optimize(1);

    g2o::OptimizableGraph::VertexSet vset;
    _vf.findVerticesScanMatching(_lastVertex, vset);
    cgm::checkCovariance(_graph, _lastVertex, vset);
    cgm::addNeighboringVertices(_graph, _lastVertex, vset, 8);
    checkHaveLaser(vset);

    std::set<g2o::OptimizableGraph::VertexSet> setOfVSet;
    _vf.findSetsOfVertices(vset, setOfVSet);

    g2o::OptimizableGraph::EdgeSet loopClosingEdges;
    for (g2o::OptimizableGraph::VertexSet myvset : setOfVSet) {
      g2o::OptimizableGraph::Vertex* closestV = _vf.findClosestVertex(myvset, _lastVertex);
      if (closestV->id() == _lastVertex->id() - 1) continue;  // this edge exists already
      if (!isMyVertex(closestV) || std::abs(_lastVertex->id() - closestV->id()) > 10) {
        // loop closure: candidates wait in the closure buffer for the vote
        std::vector<g2o::SE2> results;
        if (_LCMatcher.scanMatchingLC(myvset, closestV, _lastVertex, results, maxScore)) {
          for (const g2o::SE2& r : results) {
            g2o::EdgeSE2* ne = newEdge(closestV, _lastVertex, r, _SMinf);
            loopClosingEdges.insert(ne);
            _candidates.push_back(ne);
            log('L', ne);
          }
        } else {
          logReject(closestV->id(), _lastVertex->id());
        }
      } else {
        // edge between close vertices: added right away
        g2o::SE2 transf;
        if (_closeMatcher.closeScanMatching(myvset, closestV, _lastVertex, &transf, maxScore)) {
          g2o::EdgeSE2* ne = newEdge(closestV, _lastVertex, transf, _SMinf);
          _graph->addEdge(ne);
          log('C', ne);
        } else {
          logReject(closestV->id(), _lastVertex->id());
        }
      }
    }
    if (loopClosingEdges.size()) addClosures(loopClosingEdges);
    checkClosures();
    updateClosures();
  }

  void optimize(int nrunnings) {
    _graph->initializeOptimization();
    _graph->optimize(nrunnings);
    for (auto& kv : _graph->vertices()) {
      g2o::VertexSE2* v = static_cast<g2o::VertexSE2*>(kv.second);
      if (g2o::RobotLaser* rl = findLaserData(v)) rl->setOdomPose(v->estimate());
    }
  }
  bool saveGraph(const char* filename) { return _graph->save(filename); }
  bool loadGraph(const char* filename) { return _graph->load(filename); }

  const std::vector<Event>& events() const { return _events; }
  void clearEvents() { _events.clear(); }

 protected:
  void firstVertex(const g2o::SE2& pose, g2o::RobotLaser* laser) {
    _lastVertex = new g2o::VertexSE2();
    _lastVertex->setEstimate(pose);
    _lastVertex->setId(idRobot() * baseId());
    _lastVertex->addUserData(laser);
    _graph->addVertex(_lastVertex);
    _firstRobotPose = _lastVertex;
    _firstRobotPose->setFixed(true);
  }
  g2o::VertexSE2* newVertex(const g2o::SE2& displacement, g2o::RobotLaser* laser) {
    g2o::VertexSE2* v = new g2o::VertexSE2();
    v->setEstimate(_lastVertex->estimate() * displacement);
    v->setId(++_runningVertexId + idRobot() * baseId());
    v->addUserData(laser);
    _graph->addVertex(v);
    return v;
  }
  g2o::EdgeSE2* newEdge(g2o::OptimizableGraph::Vertex* from, g2o::OptimizableGraph::Vertex* to,
                        const g2o::SE2& z, const Eigen::Matrix3d& info) {
    g2o::EdgeSE2* e = new g2o::EdgeSE2();
    e->setId(++_runningEdgeId + idRobot() * baseId());
    e->vertices()[0] = from;
    e->vertices()[1] = to;
    e->setMeasurement(z);
    e->setInformation(info);
    // candidates are ordered by creation inside the closure buffer (g2o: by pointer)
    e->setSerial(1000000000LL + _runningEdgeId);
    return e;
  }
  g2o::RobotLaser* findLaserData(g2o::OptimizableGraph::Vertex* v) {
    return dynamic_cast<g2o::RobotLaser*>(v->userData());
  }
  void checkHaveLaser(g2o::OptimizableGraph::VertexSet& vset) {
    const g2o::OptimizableGraph::VertexSet tmp = vset;
    for (g2o::HyperGraph::Vertex* v : tmp)
      if (!findLaserData(static_cast<g2o::OptimizableGraph::Vertex*>(v))) vset.erase(v);
  }
  void addClosures(g2o::OptimizableGraph::EdgeSet loopClosingEdges) {
    _closures.addEdgeSet(loopClosingEdges);
    _closures.addVertex(_lastVertex);
  }
  void checkClosures() {
    if (!_closures.checkList(windowLoopClosure)) return;
    lcc.init(_closures.vertices(), _closures.edgeSet(), inlierThreshold);
    lcc.check();
    if (lcc.inliers() < minInliers) return;
    std::vector<g2o::OptimizableGraph::Edge*> accepted;
    for (auto& kv : lcc.closures())
      if (kv.second < inlierThreshold && !_graph->edges().count(kv.first)) accepted.push_back(kv.first);
    for (g2o::OptimizableGraph::Edge* e : accepted) {
      // addEdge gives the edge its insertion serial, which is the key of the buffer's edge set:
      // take it out and put it back (the reference keeps accepted edges in the buffer as well)
      _closures.removeEdge(e);
      _graph->addEdge(e);
      _closures.addEdge(e);
      log('A', static_cast<g2o::EdgeSE2*>(e));
    }
  }
  void updateClosures() { _closures.updateList(windowLoopClosure); }
  void log(char kind, g2o::EdgeSE2* e) {
    const Event ev = {kind, e->vertex(0)->id(), e->vertex(1)->id(), e->measurement()};
    _events.push_back(ev);
  }
  void logReject(int from, int to) {
    const Event ev = {'R', from, to, g2o::SE2()};
    _events.push_back(ev);
  }

  g2o::SparseOptimizer* _graph;
  int _idRobot, _baseId, _runningVertexId, _runningEdgeId;
  g2o::VertexSE2* _firstRobotPose;
  g2o::VertexSE2* _lastVertex;
  g2o::SE2 _lastOdom;
  VerticesFinder _vf;
  ScanMatcher _closeMatcher, _LCMatcher;
  ClosureBuffer _closures;
  int windowLoopClosure;
  double maxScore, inlierThreshold;
  int minInliers;
  LoopClosureChecker lcc;
  Eigen::Matrix3d _odominf, _SMinf;
  std::vector<Event> _events;
  std::vector<g2o::EdgeSE2*> _candidates;
};

#endif
