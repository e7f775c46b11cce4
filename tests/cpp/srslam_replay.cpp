// Offline replay of the reference's single-robot keyframe loop (src/srslam.cpp:190-221) through
// the GraphSLAM mirror (include/cgm/graph_slam.hpp) on keyframes extracted from a reference bag
// (tools/extract_bag_keyframes.py): addDataSM -> findConstraints -> optimize(5) per keyframe, with
// the reference's default parameters (srslam.cpp:77-99: resolution 0.025, kernelRadius 0.2,
// windowLoopClosure 10, maxScore 0.15, inlierThreshold 2, minInliers 7). Prints every decision and
// the final estimates; tests/test_replay_gpu.py compares them with the CPU oracle pipeline.
//
// usage:  srslam_replay keyframes.txt [graph.g2o [id_robot]]
// input:  n_beams first_angle step max_range laser_x laser_y laser_th min_inliers
//         then one line per keyframe: odom_x odom_y odom_th r_1 ... r_n
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "cgm/graph_slam.hpp"

using namespace g2o;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream f(argv[1]);
  if (!f) return 2;
  int nb, min_inliers;
  double first, step, maxr, lx, ly, lth;
  f >> nb >> first >> step >> maxr >> lx >> ly >> lth >> min_inliers;
  GraphSLAM gslam;
  gslam.setIdRobot(argc > 3 ? std::atoi(argv[3]) : 0);   // vertex ids = id_robot * 10000 + k (graph_slam.cpp:384-386)
  gslam.setBaseId(10000);
  gslam.init(0.025, 0.2, 10, 0.15, 2.0, min_inliers);
  printf("BEGIN\n");
  int k = 0;
  double ox, oy, oth;
  while (f >> ox >> oy >> oth) {
    std::vector<double> r(nb);
    for (int i = 0; i < nb; ++i) f >> r[i];
    RobotLaser* rl = new RobotLaser();
    LaserParameters lp(0, nb, first, step, maxr, 0.1, 0);
    lp.laserPose = SE2(lx, ly, lth);
    rl->setLaserParams(lp);
    rl->setRanges(r);
    const SE2 odom(ox, oy, oth);
    if (k == 0) {
      gslam.setInitialData(odom, rl);
    } else {
      gslam.addDataSM(odom, rl);
      gslam.findConstraints();
      gslam.optimize(5);
    }
    printf("K %d %d\n", k, gslam.lastVertex()->id());
    for (const GraphSLAM::Event& ev : gslam.events())
      printf("EV %c %d %d %.17g %.17g %.17g\n", ev.kind, ev.from, ev.to, ev.measurement.translation().x(),
             ev.measurement.translation().y(), ev.measurement.rotation().angle());
    gslam.clearEvents();
    ++k;
  }
  for (auto& kv : gslam.graph()->vertices()) {
    const VertexSE2* v = static_cast<const VertexSE2*>(kv.second);
    printf("P %d %.17g %.17g %.17g\n", v->id(), v->estimate().translation().x(),
           v->estimate().translation().y(), v->estimate().rotation().angle());
  }
  printf("EDGES %zu\n", gslam.graph()->edges().size());
  if (argc > 2) printf("SAVE %d\n", gslam.saveGraph(argv[2]) ? 1 : 0);  // g2o text, graph_slam.cpp:620-623
  printf("END\n");
  return 0;
}
