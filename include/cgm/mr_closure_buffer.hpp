// Per-peer sliding windows of inter-robot closures: host mirror of MRClosureBuffer
// (src/mrslam/mr_closure_buffer.{h,cpp}:30-118). MRGraphSLAM keeps two of them
// (mr_graph_slam.h:64-65): the peer vertices already matched against this robot's map, with their
// candidate edges (interRobotClosures), and the peer vertices still waiting for a match
// (interRobotVertices). A peer's window disappears with its last vertex.
#ifndef CGM_MR_CLOSURE_BUFFER_HPP
#define CGM_MR_CLOSURE_BUFFER_HPP

#include <map>

#include "slam_frontend.hpp"

struct MRClosureBuffer {
  MRClosureBuffer() {}
  // the windows are owned here (the reference copies the raw pointers around and never frees them)
  MRClosureBuffer(const MRClosureBuffer& o) { copy_from(o); }
  MRClosureBuffer& operator=(const MRClosureBuffer& o) {
    if (this != &o) {
      clear();
      copy_from(o);
    }
    return *this;
  }
  ~MRClosureBuffer() { clear(); }

  ClosureBuffer* findClosuresRobot(int robotId) {
    std::map<int, ClosureBuffer*>::iterator it = mrClosures.find(robotId);
    return it == mrClosures.end() ? 0 : it->second;
  }
  // vertices and candidate edges of `closures` join the peer's window (a vertex that is already
  // there is listed once more with age 0, as ClosureBuffer::addVertex does)
  void insert(ClosureBuffer& closures, int robotId) {
    ClosureBuffer* previousBuffer = findClosuresRobot(robotId);
    if (!previousBuffer) {
      mrClosures.insert(std::make_pair(robotId, new ClosureBuffer(closures)));
      return;
    }
    for (auto& kv : closures.vertices())
      previousBuffer->addVertex(static_cast<g2o::OptimizableGraph::Vertex*>(kv.second));
    for (g2o::HyperGraph::Edge* e : closures.edgeSet())
      previousBuffer->addEdge(static_cast<g2o::OptimizableGraph::Edge*>(e));
  }
  void remove(ClosureBuffer& closures, int robotId) {
    ClosureBuffer* previousBuffer = findClosuresRobot(robotId);
    if (!previousBuffer) return;
    for (auto& kv : closures.vertices())
      previousBuffer->removeVertex(static_cast<g2o::OptimizableGraph::Vertex*>(kv.second));
    for (g2o::HyperGraph::Edge* e : closures.edgeSet())
      previousBuffer->removeEdge(static_cast<g2o::OptimizableGraph::Edge*>(e));
    if (previousBuffer->vertices().empty()) drop(robotId);
  }
  // one keyframe passes for every peer
  void update(int windowSize) {
    const std::map<int, ClosureBuffer*> tmp = mrClosures;
    for (auto& kv : tmp) {
      kv.second->updateList(windowSize);
      if (kv.second->vertices().empty()) drop(kv.first);
    }
  }
  unsigned int size() const { return static_cast<unsigned int>(mrClosures.size()); }

  std::map<int, ClosureBuffer*> mrClosures;

 private:
  void drop(int robotId) {
    std::map<int, ClosureBuffer*>::iterator it = mrClosures.find(robotId);
    if (it == mrClosures.end()) return;
    delete it->second;
    mrClosures.erase(it);
  }
  void clear() {
    for (auto& kv : mrClosures) delete kv.second;
    mrClosures.clear();
  }
  void copy_from(const MRClosureBuffer& o) {
    for (auto& kv : o.mrClosures) mrClosures.insert(std::make_pair(kv.first, new ClosureBuffer(*kv.second)));
  }
};

#endif
