// Two robots in one process exchanging condensed graphs through the inter-robot wire format
// (src/cg_mrslam.cpp:236-259 + src/mrslam/graph_comm.cpp:126-154 drive the same calls over UDP):
//   robot B, which closed loops with some of robot A's vertices, asks A about them
//     (MRGraphSLAM::constructCondensedGraphMessage -> toCharArray -> createMsgfromCharArray ->
//      addInterRobotData on A: insertOutClosure + computeCondensedGraph on the GPU),
//   A answers with the star it computed (the same path in the other direction: B inserts the
//     star as level-0 edges between its copies of A's vertices),
//   B optimises its graph, star included.
// Input (one file):  V robot id x y th fixed | E robot i j dx dy dth I11 I12 I13 I22 I23 I33 |
//                    WANT robot peer k id_1 ... id_k   (robot's "in" closures of that peer)
// Output: the datagram sizes, the star as B received it, B's estimates after optimize(5).
// With a second argument "graph" the robots use GraphMessages instead (constructGraphMessage, the
// alternative the reference keeps next to the condensed graphs, graph_comm.cpp:147): A answers
// with its whole own graph, B adds A's vertices and edges to its graph.
#include <cstdio>
#include <fstream>
#include <sstream>

#include "mrslam/mr_graph_slam.h"

#include "driver_out.h"

using namespace g2o;

struct Robot : public MRGraphSLAM {
  CondensedGraphBuffer& buffer() { return condensedGraphs; }
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
  std::string line, tag;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    if (!(ss >> tag)) continue;
    int r;
    ss >> r;
    SparseOptimizer* g = robots[r].graph();
    if (tag == "V") {
      int id, fixed;
      double x, y, th;
      ss >> id >> x >> y >> th >> fixed;
      VertexSE2* v = new VertexSE2();
      v->setId(id);
      v->setEstimate(SE2(x, y, th));
      v->setFixed(fixed != 0);
      g->addVertex(v);
    } else if (tag == "E") {
      int a, b;
      double x, y, th, w[6];
      ss >> a >> b >> x >> y >> th;
      for (int q = 0; q < 6; ++q) ss >> w[q];
      EdgeSE2* e = new EdgeSE2();
      e->vertices()[0] = g->vertex(a);
      e->vertices()[1] = g->vertex(b);
      e->setMeasurement(SE2(x, y, th));
      Eigen::Matrix3d m;
      m(0, 0) = w[0]; m(0, 1) = m(1, 0) = w[1]; m(0, 2) = m(2, 0) = w[2];
      m(1, 1) = w[3]; m(1, 2) = m(2, 1) = w[4]; m(2, 2) = w[5];
      e->setInformation(m);
      g->addEdge(e);
    } else if (tag == "WANT") {
      int peer, k;
      ss >> peer >> k;
      OptimizableGraph::VertexIDMap want;
      for (int i = 0; i < k; ++i) {
        int id;
        ss >> id;
        want.insert(std::make_pair(id, g->vertex(id)));
      }
      robots[r].buffer().insertInClosure(peer, want);
    }
  }
  const bool whole_graph = argc > 2 && std::string(argv[2]) == "graph";
  printf("BEGIN\n");
  static char buf[MAX_LENGTH_MSG];
  // one exchange: from -> to
  auto send = [&](int from, int to) -> bool {
    size_t n = 0;
    if (whole_graph) {
      GraphMessage* out = robots[from].constructGraphMessage(to);
      if (!out) {
        printf("MSG %d %d none\n", from, to);
        return false;
      }
      char* end = out->toCharArray(buf, MAX_LENGTH_MSG);
      n = end ? static_cast<size_t>(end - buf) : 0;
      printf("MSG %d %d %zu closures %zu edges %zu vertices %zu\n", from, to, n, out->closures.size(),
             out->edgeVector.size(), out->vertexVector.size());
      delete out;
    } else {
      CondensedGraphMessage* out = robots[from].constructCondensedGraphMessage(to);
      if (!out) {
        printf("MSG %d %d none\n", from, to);
        return false;
      }
      char* end = out->toCharArray(buf, MAX_LENGTH_MSG);
      n = end ? static_cast<size_t>(end - buf) : 0;
      printf("MSG %d %d %zu closures %zu edges %zu\n", from, to, n, out->closures.size(), out->edgeVector.size());
      delete out;
    }
    if (!n) return false;
    RobotMessage* in = robots[to].createMsgfromCharArray(buf, n);
    if (!in) return false;
    StampedRobotMessage stamped = {0, in};
    robots[to].addInterRobotData(stamped);
    delete in;
    return true;
  };
  send(1, 0);   // B asks
  send(0, 1);   // A answers with its star
  OptimizableGraph::EdgeSet star = robots[1].buffer().inCondensedGraph(0);
  printf("STAR %zu %zu\n", star.size(), robots[1].graph()->edges().size());
  for (HyperGraph::Edge* he : star) {
    EdgeSE2* e = static_cast<EdgeSE2*>(he);
    printf("C %d %d %.17g %.17g %.17g", e->vertex(0)->id(), e->vertex(1)->id(), e->measurement().translation().x(),
           e->measurement().translation().y(), e->measurement().rotation().angle());
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) printf(" %.17g", e->information()(i, j));
    printf("\n");
  }
  robots[1].optimize(5);
  for (auto& kv : robots[1].graph()->vertices()) {
    const VertexSE2* v = static_cast<const VertexSE2*>(kv.second);
    printf("P %d %.17g %.17g %.17g\n", v->id(), v->estimate().translation().x(), v->estimate().translation().y(),
           v->estimate().rotation().angle());
  }
  // a second round: A's star is already known to B (replaced, not duplicated); B has nothing new
  send(1, 0);
  send(0, 1);
  printf("AGAIN %zu %zu %zu\n", robots[1].buffer().inCondensedGraph(0).size(), robots[1].graph()->edges().size(),
         robots[0].graph()->edges().size());
  // if A asked B about vertices as well, B's star has reached A by now: A optimises with it
  OptimizableGraph::EdgeSet back = robots[0].buffer().inCondensedGraph(1);
  printf("BACK %zu\n", back.size());
  for (HyperGraph::Edge* he : back) {
    EdgeSE2* e = static_cast<EdgeSE2*>(he);
    printf("D %d %d %.17g %.17g %.17g", e->vertex(0)->id(), e->vertex(1)->id(), e->measurement().translation().x(),
           e->measurement().translation().y(), e->measurement().rotation().angle());
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) printf(" %.17g", e->information()(i, j));
    printf("\n");
  }
  if (!back.empty()) {
    robots[0].optimize(5);
    for (auto& kv : robots[0].graph()->vertices()) {
      const VertexSE2* v = static_cast<const VertexSE2*>(kv.second);
      printf("Q %d %.17g %.17g %.17g\n", v->id(), v->estimate().translation().x(), v->estimate().translation().y(),
             v->estimate().rotation().angle());
    }
  }
  printf("END\n");
  return 0;
}
