// A peer's keyframe arrives as a ComboMessage (src/mrslam/mr_graph_slam.cpp:118-252, 254-329):
// robot 1 builds the message from its last vertex and scan, it crosses the float32 wire format,
// robot 0 matches the scan against its map around a reference vertex (ScanMatcher::globalMatching
// on the GPU); a match becomes a candidate inter-robot closure, a miss waits and is retried by
// findInterRobotConstraints against the map around the last vertex; the per-peer window votes and
// the accepted closure enters the graph, after which robot 0 asks robot 1 about that vertex.
// Input: V robot id x y th fixed n_beams first_angle step max_range r_1 ... r_n   (ids consecutive)
//        RUN ref_id window min_inliers max_score
#include <cstdio>
#include <fstream>
#include <sstream>

#include "mrslam/mr_graph_slam.h"

#include "driver_out.h"

using namespace g2o;

struct Robot : public MRGraphSLAM {
  void setLast(VertexSE2* v) { _lastVertex = v; }
  CondensedGraphBuffer& buffer() { return condensedGraphs; }
  int peerInGraph() {  // vertices of other robots in this robot's graph
    int n = 0;
    for (auto& kv : graph()->vertices()) n += !isMyVertex(static_cast<OptimizableGraph::Vertex*>(kv.second));
    return n;
  }
  void dump(const char* tag, int peer) {
    ClosureBuffer* c = interRobotClosures.findClosuresRobot(peer);
    ClosureBuffer* w = interRobotVertices.findClosuresRobot(peer);
    printf("%s closures %zu waiting %zu graph_edges %zu peer_in_graph %d\n", tag, c ? c->vertices().size() : 0,
           w ? w->vertices().size() : 0, graph()->edges().size(), peerInGraph());
    if (c)
      for (HyperGraph::Edge* he : c->edgeSet()) {
        EdgeSE2* e = static_cast<EdgeSE2*>(he);
        printf("CAND %d %d %.17g %.17g %.17g\n", e->vertex(0)->id(), e->vertex(1)->id(),
               e->measurement().translation().x(), e->measurement().translation().y(),
               e->measurement().rotation().angle());
      }
  }
};

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream f(argv[1]);
  if (!f) return 2;
  Robot robots[2];
  for (int r = 0; r < 2; ++r) {
    robots[r].setIdRobot(r);
    robots[r].setBaseId(10000);
    robots[r].init(0.025, 0.2, 10, 0.15, 2.0, 7);
  }
  VertexSE2* last[2] = {0, 0};
  int ref_id = -1, window = 2, min_inliers = 1;
  double max_score = 0.3;
  std::string line, tag;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    if (!(ss >> tag)) continue;
    if (tag == "V") {
      int r, id, fixed, nb;
      double x, y, th, first, step, maxr;
      ss >> r >> id >> x >> y >> th >> fixed >> nb >> first >> step >> maxr;
      VertexSE2* v = new VertexSE2();
      v->setId(id);
      v->setEstimate(SE2(x, y, th));
      v->setFixed(fixed != 0);
      std::vector<double> rr(nb);
      for (int i = 0; i < nb; ++i) ss >> rr[i];
      RobotLaser* rl = new RobotLaser();
      LaserParameters lp(0, nb, first, step, maxr, 0.1, 0);
      lp.laserPose = SE2(0.05, 0.0, 0.0);
      rl->setLaserParams(lp);
      rl->setRanges(rr);
      v->setUserData(rl);
      robots[r].graph()->addVertex(v);
      last[r] = v;
    } else if (tag == "RUN") {
      ss >> ref_id >> window >> min_inliers >> max_score;
    }
  }
  for (int r = 0; r < 2; ++r) robots[r].setLast(last[r]);
  robots[0].setInterRobotClosureParams(max_score, min_inliers, window);
  printf("BEGIN\n");
  static char buf[MAX_LENGTH_MSG];
  ComboMessage* out = robots[1].constructComboMessage();
  char* end = out->toCharArray(buf, MAX_LENGTH_MSG);
  const size_t n = end ? static_cast<size_t>(end - buf) : 0;
  printf("MSG %zu vertices %zu readings %zu node %d\n", n, out->vertexVector.size(), out->readings.size(), out->nodeId);
  delete out;
  RobotMessage* in = robots[0].createMsgfromCharArray(buf, n);
  if (!in) return 3;
  StampedRobotMessage stamped = {robots[0].graph()->vertex(ref_id), in};
  robots[0].addInterRobotData(stamped);
  delete in;
  robots[0].dump("ARRIVAL", 1);
  for (int round = 0; round < 3; ++round) {
    robots[0].findInterRobotConstraints();
    robots[0].dump("ROUND", 1);
  }
  CondensedGraphMessage* ask = robots[0].constructCondensedGraphMessage(1);
  printf("ASK %zu", ask ? ask->closures.size() : 0);
  if (ask)
    for (int id : ask->closures) printf(" %d", id);
  printf("\n");
  delete ask;
  // the same keyframe again: every vertex is known by now, nothing new is matched
  out = robots[1].constructComboMessage();
  end = out->toCharArray(buf, MAX_LENGTH_MSG);
  delete out;
  in = robots[0].createMsgfromCharArray(buf, static_cast<size_t>(end - buf));
  StampedRobotMessage again = {robots[0].graph()->vertex(ref_id), in};
  robots[0].addInterRobotData(again);
  delete in;
  robots[0].dump("REPEAT", 1);
  printf("END\n");
  return 0;
}
