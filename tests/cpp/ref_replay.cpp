// Offline replay of the reference's single-robot main loop (src/srslam.cpp:130-139,190-221) driving
// the REFERENCE'S OWN GraphSLAM (src/slam/graph_slam.cpp and everything under it, compiled verbatim
// from /root/reference by oracle/Makefile, target `frontend`). The same driver and the same
// reference sources are linked twice:
//   oracle/_ref/ref_replay_gpu  -- against include/cgm/chargrid.hpp + include/g2o_compat +
//                                  libcgmrslam_b200.so: the reference's front-end over the CUDA paths;
//   oracle/_ref/ref_replay_cpu  -- against the reference's own chargrid.cpp (CPU matcher) and the
//                                  CPU oracle solver (oracle/pgo_oracle_c.cpp behind the same ABI).
// tests/test_ref_frontend.py runs both on keyframes extracted from the reference's bags and compares
// every decision (which edges enter the graph, which closures are buffered / accepted: vertex
// indices exact) and every estimate (1e-6).
//
// The pipeline is chaotic at the float-ulp level (the matcher's window bounds are floats:
// scan_matcher.cpp:147-150 turns a 1e-13 difference of an estimate into a 7e-9 difference of a
// measurement now and then, and such differences eventually flip a discrete decision), so two
// different solvers cannot be compared free-running over ~900 keyframes. They are compared in
// LOCKSTEP instead, at the granularity of the solver calls (g2o::trace in g2o_compat.hpp): the GPU
// build records the result of every optimize / computeMarginals / ... (--dump), the CPU build
// follows (--follow): it computes every call from the same inputs, the trace notes how far its
// result is from the GPU's, and it carries on with the GPU's numbers (line "D k worst-so-far"; the
// totals per kind of call are on the TRACE line).
//
// usage:  ref_replay keyframes.txt [graph.g2o [id_robot [n_keyframes [--dump|--follow states.bin]]]]
// input:  n_beams first_angle step max_range laser_x laser_y laser_th min_inliers
//         then one line per keyframe: odom_x odom_y odom_th r_1 ... r_n
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <set>
#include <sstream>
#include <vector>

#include "slam/graph_slam.h"

#include "driver_out.h"

using namespace g2o;

namespace {

struct EdgeKey {
  int from, to;
  double x, y, th, w;
  bool operator<(const EdgeKey& o) const {
    if (from != o.from) return from < o.from;
    if (to != o.to) return to < o.to;
    if (x != o.x) return x < o.x;
    if (y != o.y) return y < o.y;
    return th < o.th;
  }
};

EdgeKey key_of(const HyperGraph::Edge* he) {
  const EdgeSE2* e = static_cast<const EdgeSE2*>(he);
  EdgeKey k = {e->vertices()[0]->id(), e->vertices()[1]->id(), e->measurement().translation().x(),
               e->measurement().translation().y(), e->measurement().rotation().angle(), e->information()(0, 0)};
  return k;
}

// the reference keeps its closure window protected; a subclass may look (no source edit)
struct Probe : public GraphSLAM {
  std::multiset<EdgeKey> buffered() {
    std::multiset<EdgeKey> s;
    for (OptimizableGraph::EdgeSet::iterator it = _closures.edgeSet().begin(); it != _closures.edgeSet().end(); ++it)
      s.insert(key_of(*it));
    return s;
  }
};

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream f(argv[1]);
  if (!f) return 2;
  int nb, min_inliers;
  double first, step, maxr, lx, ly, lth;
  f >> nb >> first >> step >> maxr >> lx >> ly >> lth >> min_inliers;
  const int id_robot = argc > 3 ? std::atoi(argv[3]) : 0;
  const int limit = argc > 4 ? std::atoi(argv[4]) : 1 << 30;
  // lockstep at the granularity of the solver calls (include/g2o_compat/g2o_compat.hpp, namespace trace)
  bool follow = false;
  if (argc > 6 && std::string(argv[5]) == "--dump" && !g2o::trace::open(0, argv[6], false)) return 2;
  if (argc > 6 && std::string(argv[5]) == "--follow") {
    if (!g2o::trace::open(0, argv[6], true)) return 2;
    follow = true;
  }
  Probe gslam;
  gslam.setIdRobot(id_robot);  // vertex ids = id_robot * 10000 + k (graph_slam.cpp:92,157)
  gslam.setBaseId(10000);
  gslam.init(0.025, 0.2, 10, 0.15, 2.0, min_inliers);  // srslam.cpp:77-99 defaults
  printf("BEGIN\n");
  std::multiset<EdgeKey> graph_before, buffer_before;
  SE2 odom_prev, currEst;
  double t_sm = 0.0, t_fc = 0.0, t_opt = 0.0;
  int k = 0;
  double ox, oy, oth;
  while (k < limit && f >> ox >> oy >> oth) {
    std::vector<double> r(nb);
    for (int i = 0; i < nb; ++i) f >> r[i];
    RobotLaser* rl = new RobotLaser();
    LaserParameters lp(0, nb, first, step, maxr, 0.1, 0);  // ros_handler.cpp:92-106
    lp.laserPose = SE2(lx, ly, lth);
    rl->setLaserParams(lp);
    rl->setRanges(r);
    const SE2 odom(ox, oy, oth);
    if (k == 0) {
      gslam.setInitialData(odom, rl);  // srslam.cpp:130-139
      currEst = gslam.lastVertex()->estimate();
    } else {
      // srslam.cpp:193-196: the estimate of the last keyframe moved by the odometry increment
      currEst *= odom_prev.inverse() * odom;
      double t0 = now_ms();
      gslam.addDataSM(currEst, rl);
      double t1 = now_ms();
      gslam.findConstraints();
      double t2 = now_ms();
      gslam.optimize(5);
      double t3 = now_ms();
      t_sm += t1 - t0;
      t_fc += t2 - t1;
      t_opt += t3 - t2;
      printf("T %d %.4f %.4f %.4f\n", k, t1 - t0, t2 - t1, t3 - t2);  // ms: addDataSM, findConstraints, optimize(5)
      currEst = gslam.lastVertex()->estimate();  // srslam.cpp:214
    }
    odom_prev = odom;
    if (follow) {
      const g2o::trace::Stream& t = g2o::trace::stats(0);
      if (t.parted_at >= 0) {
        printf("D %d -1\n", k);  // the leader made a different solver call here: the runs have parted
        return 4;
      }
      // the follower's own results against the leader's, every solver call so far: estimates and
      // star measurements (absolute), covariance / information blocks (relative)
      double worst = 0.0;
      for (int q = 1; q < g2o::trace::kKinds; ++q) worst = std::max(worst, t.worst[q]);
      printf("D %d %.3e\n", k, worst);
    }
    const SE2 est = gslam.lastVertex()->estimate();
    printf("K %d %d %.17g %.17g %.17g\n", k, gslam.lastVertex()->id(), est.translation().x(),
           est.translation().y(), est.rotation().angle());
    std::multiset<EdgeKey> graph_now, buffer_now = gslam.buffered();
    for (HyperGraph::EdgeSet::const_iterator it = gslam.graph()->edges().begin(); it != gslam.graph()->edges().end(); ++it)
      graph_now.insert(key_of(*it));
    std::vector<EdgeKey> fresh;
    std::set_difference(graph_now.begin(), graph_now.end(), graph_before.begin(), graph_before.end(),
                        std::back_inserter(fresh));
    for (size_t i = 0; i < fresh.size(); ++i)
      printf("E %d %d %.17g %.17g %.17g %.17g\n", fresh[i].from, fresh[i].to, fresh[i].x, fresh[i].y,
             fresh[i].th, fresh[i].w);
    fresh.clear();
    std::set_difference(buffer_now.begin(), buffer_now.end(), buffer_before.begin(), buffer_before.end(),
                        std::back_inserter(fresh));
    for (size_t i = 0; i < fresh.size(); ++i)
      printf("C %d %d %.17g %.17g %.17g\n", fresh[i].from, fresh[i].to, fresh[i].x, fresh[i].y, fresh[i].th);
    graph_before.swap(graph_now);
    buffer_before.swap(buffer_now);
    ++k;
  }
  for (HyperGraph::VertexIDMap::const_iterator it = gslam.graph()->vertices().begin();
       it != gslam.graph()->vertices().end(); ++it) {
    const VertexSE2* v = static_cast<const VertexSE2*>(it->second);
    printf("P %d %.17g %.17g %.17g\n", v->id(), v->estimate().translation().x(),
           v->estimate().translation().y(), v->estimate().rotation().angle());
  }
  printf("EDGES %zu\n", gslam.graph()->edges().size());
  {
    const g2o::trace::Stream& t = g2o::trace::stats(0);
    printf("TRACE optimize %lld %.3e initial_guess %lld %.3e marginals %lld %.3e star_meas %lld %.3e star_info %lld %.3e\n",
           t.calls[1], t.worst[1], t.calls[2], t.worst[2], t.calls[3], t.worst[3], t.calls[4], t.worst[4], t.calls[5],
           t.worst[5]);
    g2o::trace::close(0);
  }
  {
    const double* w = g2o::trace::wall_ms();
    printf("SOLVER_MS set_graph %.3f %.0f upload %.3f %.0f iterate %.3f %.0f marginals %.3f %.0f initial_guess %.3f %.0f\n", w[0],
           w[6], w[1], w[7], w[2], w[8], w[3], w[9], w[4], w[10]);
  }
  printf("TIMES_MS addDataSM %.3f findConstraints %.3f optimize %.3f keyframes %d\n", t_sm, t_fc, t_opt, k);
  if (argc > 2 && std::string(argv[2]) != "-") printf("SAVE %d\n", gslam.saveGraph(argv[2]) ? 1 : 0);
  printf("END\n");
  return 0;
}
