// Multi-robot graph SLAM: host mirror of MRGraphSLAM (src/mrslam/mr_graph_slam.{h,cpp}:32-746)
// -- SURVEY 8f row 4 -- on top of the GraphSLAM mirror, the inter-robot message classes
// (msg_factory.hpp), the per-peer closure windows (mr_closure_buffer.hpp) and the condensed-graph
// buffer (condensed_graph.hpp). Same member names and argument meaning as the reference; the
// graph mutex of the reference is the caller's business here (one thread drives a robot).
//
// What reaches the GPU from here: ScanMatcher::globalMatching (hierarchical correlative search of
// a peer's scan against the local map) for every ComboMessage and every still-unmatched peer
// vertex, LoopClosureChecker votes over the per-peer windows, and -- through
// CondensedGraphBuffer::computeCondensedGraph -- one optimize(1) + labelEdges per star a peer asks
// for. Everything a peer learns about this robot crosses the float32 wire format.
#ifndef CGM_MR_GRAPH_SLAM_HPP
#define CGM_MR_GRAPH_SLAM_HPP

#include <map>
#include <utility>

#include "condensed_graph.hpp"
#include "graph_slam.hpp"
#include "mr_closure_buffer.hpp"
#include "msg_factory.hpp"

struct StampedRobotMessage {
  g2o::OptimizableGraph::Vertex* refVertex;  // current vertex when the message was received
  RobotMessage* msg;
};

class MRGraphSLAM : public GraphSLAM {
 public:
  MRGraphSLAM()
      : condensedGraphs(GraphSLAM::graph()), maxScoreMR(GraphSLAM::maxScore), minInliersMR(GraphSLAM::minInliers),
        windowMRLoopClosure(GraphSLAM::windowLoopClosure), detectRobotInRange(false) {
    factory.registerMessageType<ComboMessage>();
    factory.registerMessageType<CondensedGraphMessage>();
    factory.registerMessageType<GraphMessage>();
    factory.registerMessageType<EdgeArrayMessage>();
    factory.registerMessageType<VertexArrayMessage>();
    factory.registerMessageType<RobotLaserMessage>();
  }
  ~MRGraphSLAM() {
    // peer vertices that never entered the graph are ours (the reference leaks them)
    for (g2o::VertexSE2* v : _peerVertices)
      if (graph()->vertex(v->id()) != v) delete v;
    for (g2o::EdgeSE2* e : _peerEdges)
      if (!graph()->edges().count(e)) delete e;
  }

  void setInterRobotClosureParams(double maxScoreMR_, int minInliersMR_, int windowMRLoopClosure_) {
    maxScoreMR = maxScoreMR_;
    minInliersMR = minInliersMR_;
    windowMRLoopClosure = windowMRLoopClosure_;
  }
  void setDetectRobotInRange(bool detectRobotInRange_) { detectRobotInRange = detectRobotInRange_; }

  // dispatch on the message's type (mr_graph_slam.cpp:485-500)
  void addInterRobotData(StampedRobotMessage vmsg) {
    if (ComboMessage* cmsg = dynamic_cast<ComboMessage*>(vmsg.msg)) {
      addInterRobotData(cmsg, vmsg.refVertex);
    } else if (CondensedGraphMessage* cgmsg = dynamic_cast<CondensedGraphMessage*>(vmsg.msg)) {
      addInterRobotData(cgmsg);
    } else if (GraphMessage* gmsg = dynamic_cast<GraphMessage*>(vmsg.msg)) {
      addInterRobotData(gmsg);
    }
  }

  // Once per keyframe: try the still-unmatched peer vertices against the map around the last
  // vertex (up to 20 predecessors), then vote over every peer's window and age the windows
  // (mr_graph_slam.cpp:254-329).
  void findInterRobotConstraints() {
    g2o::OptimizableGraph::VertexSet referenceVset;
    g2o::VertexSE2* referenceVertex = lastVertex();
    referenceVset.insert(referenceVertex);
    const int gap = 20;
    for (int j = 1; j <= gap; j++) {
      g2o::VertexSE2* vj = dynamic_cast<g2o::VertexSE2*>(graph()->vertex(referenceVertex->id() - j));
      if (!vj) break;
      referenceVset.insert(vj);
    }
    MRClosureBuffer tmp = interRobotVertices;
    for (auto& kv : tmp.mrClosures) {
      const int robotId = kv.first;
      g2o::OptimizableGraph::VertexIDMap vertices = kv.second->vertices();
      for (auto& iv : vertices) {
        g2o::VertexSE2* v = static_cast<g2o::VertexSE2*>(iv.second);
        g2o::SE2 transf;
        if (!_LCMatcher.globalMatching(referenceVset, referenceVertex, v, &transf, maxScoreMR)) continue;
        if (detectRobotInRange) {
          double score;
          g2o::OptimizableGraph::VertexSet vset;
          vset.insert(v);
          if (!_LCMatcher.verifyMatching(referenceVset, referenceVertex, vset, v, transf, &score)) continue;
        }
        // matched: it moves from interRobotVertices to interRobotClosures
        ClosureBuffer closure;
        closure.addVertex(v);
        closure.addEdge(interRobotEdge(referenceVertex, v, transf));
        interRobotClosures.insert(closure, robotId);
        interRobotVertices.remove(closure, robotId);
      }
    }
    checkInterRobotClosures();
    updateInterRobotClosures();
    interRobotVertices.update(windowMRLoopClosure);
  }

  // the last vertex, up to four predecessors and the last scan (mr_graph_slam.cpp:564-605)
  ComboMessage* constructComboMessage() {
    ComboMessage* cmsg = dynamic_cast<ComboMessage*>(factory.constructMessage(ComboMessage::_type()));
    cmsg->setRobotId(idRobot());
    const int nVertices = 5;
    g2o::OptimizableGraph::VertexIDMap vertices;
    vertices.insert(std::make_pair(lastVertex()->id(), lastVertex()));
    for (int i = 1; i < nVertices; i++) {
      g2o::VertexSE2* v = dynamic_cast<g2o::VertexSE2*>(graph()->vertex(lastVertex()->id() - i));
      if (!v) break;
      vertices.insert(std::make_pair(v->id(), v));
    }
    fillVertices(cmsg, vertices);
    if (g2o::RobotLaser* robotLaser = findLaserData(lastVertex())) {
      cmsg->readings = robotLaser->ranges();
      cmsg->nodeId = lastVertex()->id();
      cmsg->minangle = robotLaser->laserParams().firstBeamAngle;
      cmsg->angleincrement = robotLaser->laserParams().angularStep;
      cmsg->maxrange = robotLaser->laserParams().maxRange;
      cmsg->accuracy = robotLaser->laserParams().accuracy;
    }
    return cmsg;
  }

  // which of the peer's vertices this robot wants a star over, and the star it computed for the
  // peer; 0 when there is neither (mr_graph_slam.cpp:607-670)
  CondensedGraphMessage* constructCondensedGraphMessage(int idRobotTo) {
    CondensedGraphMessage* gmsg =
        dynamic_cast<CondensedGraphMessage*>(factory.constructMessage(CondensedGraphMessage::_type()));
    gmsg->setRobotId(idRobot());
    g2o::OptimizableGraph::VertexIDMap* inClosuresRobot = condensedGraphs.inClosures(idRobotTo);
    if (inClosuresRobot)
      for (auto& kv : *inClosuresRobot) gmsg->closures.push_back(kv.first);
    g2o::OptimizableGraph::EdgeSet subgraphRobot = condensedGraphs.outCondensedGraph(idRobotTo);
    const bool have_edges = fillEdges(gmsg, subgraphRobot);
    if (inClosuresRobot || have_edges) return gmsg;
    delete gmsg;
    return 0;
  }

  // the non-condensed alternative: once the peer has asked for vertices, this robot's whole own
  // graph (mr_graph_slam.cpp:672-739)
  GraphMessage* constructGraphMessage(int idRobotTo) {
    GraphMessage* gmsg = dynamic_cast<GraphMessage*>(factory.constructMessage(GraphMessage::_type()));
    gmsg->setRobotId(idRobot());
    g2o::OptimizableGraph::VertexIDMap* inClosuresRobot = condensedGraphs.inClosures(idRobotTo);
    if (inClosuresRobot)
      for (auto& kv : *inClosuresRobot) gmsg->closures.push_back(kv.first);
    g2o::OptimizableGraph::VertexIDMap* outClosuresRobot = condensedGraphs.outClosures(idRobotTo);
    if (outClosuresRobot) {
      g2o::OptimizableGraph::EdgeSet myOwnEdges = condensedGraphs.getMyEdges();
      fillEdges(gmsg, myOwnEdges);
      g2o::OptimizableGraph::VertexIDMap myOwnVertices;
      for (auto& kv : graph()->vertices())
        if (isMyVertex(static_cast<g2o::OptimizableGraph::Vertex*>(kv.second))) myOwnVertices.insert(kv);
      fillVertices(gmsg, myOwnVertices);
    }
    if (inClosuresRobot || outClosuresRobot) return gmsg;
    delete gmsg;
    return 0;
  }

  // 0 for a datagram that is not exactly one known message
  RobotMessage* createMsgfromCharArray(const char* buffer, size_t size) { return factory.fromCharArray(buffer, size); }

 protected:
  // A peer's keyframe: its last vertices' estimates and its last scan. Known vertices get their
  // estimate refreshed; the new one (with the scan) is matched against the map around refVertex
  // (10 vertices either side); a match becomes a candidate closure of that peer, a miss waits in
  // interRobotVertices (mr_graph_slam.cpp:118-252).
  void addInterRobotData(ComboMessage* cmsg, g2o::OptimizableGraph::Vertex* refVertex) {
    g2o::OptimizableGraph::VertexSet vset;
    for (size_t i = 0; i < cmsg->vertexVector.size(); i++) {
      const int vID = cmsg->vertexVector[i].id;
      const g2o::SE2 vest(cmsg->vertexVector[i].estimate[0], cmsg->vertexVector[i].estimate[1],
                          cmsg->vertexVector[i].estimate[2]);
      if (dynamic_cast<g2o::VertexSE2*>(graph()->vertex(vID))) continue;  // already in the graph
      bool known = false;
      MRClosureBuffer* lists[2] = {&interRobotClosures, &interRobotVertices};
      for (int k = 0; k < 2 && !known; ++k)
        if (ClosureBuffer* cb = lists[k]->findClosuresRobot(cmsg->robotId()))
          if (g2o::OptimizableGraph::Vertex* v2 = cb->findVertex(vID)) {
            g2o::VertexSE2* vse2 = static_cast<g2o::VertexSE2*>(v2);
            vse2->setEstimate(vest);
            vset.insert(vse2);
            known = true;
          }
      if (known) continue;
      if (vID == cmsg->nodeId) {  // the new vertex, with its scan
        g2o::VertexSE2* v = new g2o::VertexSE2;
        v->setId(cmsg->nodeId);
        // (the peer's max range is replaced by 8 m, mr_graph_slam.cpp:161)
        g2o::LaserParameters lparams(0, static_cast<int>(cmsg->readings.size()), cmsg->minangle,
                                     cmsg->angleincrement, 8.0, cmsg->accuracy, 0);
        g2o::RobotLaser* robotlaser = new g2o::RobotLaser;
        robotlaser->setLaserParams(lparams);
        robotlaser->setRanges(cmsg->readings);
        v->setUserData(robotlaser);
        v->setEstimate(vest);
        robotlaser->setOdomPose(vest);
        _peerVertices.push_back(v);
        vset.insert(v);
      }
    }
    g2o::OptimizableGraph::VertexSet referenceVset;
    g2o::VertexSE2* referenceVertex = static_cast<g2o::VertexSE2*>(refVertex);
    referenceVset.insert(referenceVertex);
    const int gap = 10;
    for (int dir = -1; dir <= 1; dir += 2)
      for (int j = 1; j <= gap; j++) {
        g2o::VertexSE2* vj = dynamic_cast<g2o::VertexSE2*>(graph()->vertex(referenceVertex->id() + dir * j));
        if (!vj) break;
        referenceVset.insert(vj);
      }
    if (vset.empty()) return;
    g2o::VertexSE2* v = 0;
    for (g2o::HyperGraph::Vertex* hv : vset)
      if (hv->id() == cmsg->nodeId) v = static_cast<g2o::VertexSE2*>(hv);
    if (!v) return;  // (the reference matches a fresh, scan-less vertex here: never a match)
    g2o::SE2 transf;
    const bool shouldIAdd = _LCMatcher.globalMatching(referenceVset, referenceVertex, vset, v, &transf, maxScoreMR);
    if (shouldIAdd) {
      if (detectRobotInRange) {
        double score;
        if (!_LCMatcher.verifyMatching(referenceVset, referenceVertex, vset, v, transf, &score)) return;
      }
      ClosureBuffer closure;
      closure.addVertex(v);
      closure.addEdge(interRobotEdge(referenceVertex, v, transf));
      interRobotClosures.insert(closure, cmsg->robotId());
    } else {
      ClosureBuffer c;
      c.addVertex(v);
      interRobotVertices.insert(c, cmsg->robotId());
    }
  }

  // the vertices the peer wants a star over (-> compute it), and the star it computed for us
  // (-> level-0 edges between vertices we know) (mr_graph_slam.cpp:331-395)
  void addInterRobotData(CondensedGraphMessage* gmsg) {
    takeClosures(gmsg, gmsg->robotId());
    takeEdges(gmsg, gmsg->robotId());
  }
  // the same with the peer's whole graph: unknown vertices are created, known foreign ones get
  // their estimate refreshed (mr_graph_slam.cpp:397-483)
  void addInterRobotData(GraphMessage* gmsg) {
    takeClosures(gmsg, gmsg->robotId());
    for (size_t i = 0; i < gmsg->vertexVector.size(); i++) {
      g2o::VertexSE2* v = static_cast<g2o::VertexSE2*>(graph()->vertex(gmsg->vertexVector[i].id));
      const g2o::SE2 est(gmsg->vertexVector[i].estimate[0], gmsg->vertexVector[i].estimate[1],
                         gmsg->vertexVector[i].estimate[2]);
      if (v) {
        if (!isMyVertex(v)) v->setEstimate(est);
      } else {
        v = new g2o::VertexSE2;
        v->setId(gmsg->vertexVector[i].id);
        v->setEstimate(est);
        graph()->addVertex(v);
      }
    }
    takeEdges(gmsg, gmsg->robotId());
  }

  // vote over every peer's window that is about to lose a vertex; the inliers enter the graph and
  // their peer vertices become "in" closures, i.e. what the next condensed-graph message asks the
  // peer about (mr_graph_slam.cpp:60-112)
  void checkInterRobotClosures() {
    for (auto& kv : interRobotClosures.mrClosures) {
      const int robotId = kv.first;
      ClosureBuffer* cb = kv.second;
      if (!cb->checkList(windowMRLoopClosure)) continue;
      lcc.init(cb->vertices(), cb->edgeSet(), inlierThreshold);
      lcc.check();
      if (lcc.inliers() < minInliersMR) continue;
      g2o::OptimizableGraph::VertexIDMap inClosures;
      std::vector<g2o::EdgeSE2*> accepted;
      for (auto& re : lcc.closures())
        if (re.second < inlierThreshold && !graph()->edges().count(re.first))
          accepted.push_back(static_cast<g2o::EdgeSE2*>(re.first));
      for (g2o::EdgeSE2* e : accepted) {
        g2o::VertexSE2* vto = static_cast<g2o::VertexSE2*>(e->vertices()[1]);
        e->setId(++_runningEdgeId + _baseId);
        g2o::OptimizableGraph::Vertex* inserted = graph()->vertex(vto->id());
        if (!inserted) {
          graph()->addVertex(vto);
        } else if (inserted != vto && !findLaserData(inserted)) {
          // known (from a graph message) but without its scan: it gets a copy of it (the reference
          // shares the pointer between the two vertices)
          if (g2o::RobotLaser* laservto = findLaserData(vto)) {
            g2o::RobotLaser* copy = new g2o::RobotLaser(*laservto);
            copy->setNext(0);
            inserted->setUserData(copy);
          }
        }
        // addEdge gives the edge its insertion serial, the key of the window's edge set: take it
        // out and put it back (as GraphSLAM::checkClosures does)
        cb->removeEdge(e);
        graph()->addEdge(e);
        cb->addEdge(e);
        inClosures.insert(std::make_pair(vto->id(), vto));
      }
      if (inClosures.size()) condensedGraphs.insertInClosure(robotId, inClosures);
    }
  }
  void updateInterRobotClosures() { interRobotClosures.update(windowMRLoopClosure); }

  // ---- shared pieces of the message handlers ----------------------------------------------------
  void fillVertices(VertexArrayMessage* vmsg, g2o::OptimizableGraph::VertexIDMap& vertices) {
    vmsg->vertexVector.resize(vertices.size());
    size_t i = 0;
    for (auto& kv : vertices) {
      const g2o::VertexSE2* v = static_cast<const g2o::VertexSE2*>(kv.second);
      vmsg->vertexVector[i].id = v->id();
      vmsg->vertexVector[i].estimate[0] = v->estimate().translation().x();
      vmsg->vertexVector[i].estimate[1] = v->estimate().translation().y();
      vmsg->vertexVector[i].estimate[2] = v->estimate().rotation().angle();
      ++i;
    }
  }
  bool fillEdges(EdgeArrayMessage* emsg, g2o::OptimizableGraph::EdgeSet& edges) {
    emsg->edgeVector.resize(edges.size());
    size_t i = 0;
    for (g2o::HyperGraph::Edge* he : edges) {
      g2o::EdgeSE2* e = static_cast<g2o::EdgeSE2*>(he);
      EdgeArrayMessage::ESE2Data& d = emsg->edgeVector[i++];
      d.idfrom = e->vertices()[0]->id();
      d.idto = e->vertices()[1]->id();
      d.estimate[0] = e->measurement().translation().x();
      d.estimate[1] = e->measurement().translation().y();
      d.estimate[2] = e->measurement().rotation().angle();
      const Eigen::Matrix3d inf = e->information();
      d.information[0] = inf(0, 0);
      d.information[1] = inf(0, 1);
      d.information[2] = inf(0, 2);
      d.information[3] = inf(1, 1);
      d.information[4] = inf(1, 2);
      d.information[5] = inf(2, 2);
    }
    return !edges.empty();
  }
  void takeClosures(ClosuresMessage* cmsg, int robotId) {
    g2o::OptimizableGraph::VertexIDMap closures;
    for (size_t i = 0; i < cmsg->closures.size(); i++)
      if (g2o::OptimizableGraph::Vertex* v = graph()->vertex(cmsg->closures[i]))
        closures.insert(std::make_pair(cmsg->closures[i], v));
    if (closures.empty()) return;
    condensedGraphs.insertOutClosure(robotId, closures);
    condensedGraphs.computeCondensedGraph(robotId);
  }
  void takeEdges(EdgeArrayMessage* emsg, int robotId) {
    g2o::OptimizableGraph::EdgeSet edges;
    long long serial = 1LL << 41;
    for (size_t i = 0; i < emsg->edgeVector.size(); i++) {
      const EdgeArrayMessage::ESE2Data& d = emsg->edgeVector[i];
      g2o::OptimizableGraph::Vertex* vfrom = graph()->vertex(d.idfrom);
      g2o::OptimizableGraph::Vertex* vto = graph()->vertex(d.idto);
      if (!vfrom || !vto) continue;
      g2o::EdgeSE2* e = new g2o::EdgeSE2;
      e->vertices()[0] = vfrom;
      e->vertices()[1] = vto;
      Eigen::Matrix3d inf;
      inf(0, 0) = d.information[0];
      inf(0, 1) = inf(1, 0) = d.information[1];
      inf(0, 2) = inf(2, 0) = d.information[2];
      inf(1, 1) = d.information[3];
      inf(1, 2) = inf(2, 1) = d.information[4];
      inf(2, 2) = d.information[5];
      e->setMeasurement(g2o::SE2(d.estimate[0], d.estimate[1], d.estimate[2]));
      e->setInformation(inf);
      e->setSerial(serial++);  // message order
      edges.insert(e);
    }
    if (edges.size()) condensedGraphs.insertEdgesFromRobot(robotId, edges);
  }
  // candidate inter-robot closure: information diag(100, 100, 1000) (mr_graph_slam.cpp:222-230)
  g2o::EdgeSE2* interRobotEdge(g2o::VertexSE2* from, g2o::VertexSE2* to, const g2o::SE2& transf) {
    g2o::EdgeSE2* ne = new g2o::EdgeSE2;
    ne->vertices()[0] = from;
    ne->vertices()[1] = to;
    ne->setMeasurement(transf);
    Eigen::Matrix3d inf = Eigen::Matrix3d::Identity();
    inf(0, 0) = inf(1, 1) = 100.0;
    inf(2, 2) = 1000.0;
    ne->setInformation(inf);
    ne->setSerial((1LL << 42) + static_cast<long long>(_peerEdges.size()));
    _peerEdges.push_back(ne);
    return ne;
  }

  CondensedGraphBuffer condensedGraphs;
  MRClosureBuffer interRobotClosures;  // already matched peer vertices + their candidate edges
  MRClosureBuffer interRobotVertices;  // peer vertices not matched yet
  double maxScoreMR;
  int minInliersMR;
  int windowMRLoopClosure;
  bool detectRobotInRange;
  MessageFactory factory;
  std::vector<g2o::VertexSE2*> _peerVertices;
  std::vector<g2o::EdgeSE2*> _peerEdges;
};

#endif
