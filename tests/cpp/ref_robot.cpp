// One robot of the multi-robot system behind a flat C interface: the REFERENCE'S OWN MRGraphSLAM
// (src/mrslam/mr_graph_slam.cpp and everything under it, compiled verbatim from /root/reference by
// oracle/Makefile, target `frontend`), driven like src/cg_mrslam.cpp:206-226 drives it, with the
// send / receive halves of src/mrslam/graph_comm.cpp:126-154,178-211 as two calls. Built twice:
//   oracle/_ref/libref_robot_gpu.so -- over include/cgm/chargrid.hpp + include/g2o_compat +
//                                      libcgmrslam_b200.so,
//   oracle/_ref/libref_robot_cpu.so -- over the reference's chargrid.cpp and the CPU oracle solver.
// tests/mr_replay.py schedules several robots (one process, or one robot per rank with the
// datagrams exchanged by a torch.distributed all-gather in place of UDP).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "mrslam/mr_graph_slam.h"

using namespace g2o;

namespace {

struct Robot : public MRGraphSLAM {
  SE2 odom_prev, currEst;
  int last_sent_vertex = -1;
  int nb = 0;
  double first = 0, step = 0, maxr = 0;
  SE2 laser_pose;

  RobotLaser* laser(const double* ranges) {
    RobotLaser* rl = new RobotLaser();
    LaserParameters lp(0, nb, first, step, maxr, 0.1, 0);  // ros_handler.cpp:92-106
    lp.laserPose = laser_pose;
    rl->setLaserParams(lp);
    rl->setRanges(std::vector<double>(ranges, ranges + nb));
    return rl;
  }
};

}  // namespace

extern "C" {

// cg_mrslam.cpp:137-166: sim modality, the first vertex sits at the ground-truth pose
void* robot_create(int id, double init_x, double init_y, double init_th, double odom_x, double odom_y,
                   double odom_th, int n_beams, double first_angle, double angular_step, double max_range,
                   double laser_x, double laser_y, double laser_th, const double* ranges, int min_inliers,
                   double max_score_mr, int min_inliers_mr, int window_mr) {
  Robot* r = new Robot();
  r->nb = n_beams;
  r->first = first_angle;
  r->step = angular_step;
  r->maxr = max_range;
  r->laser_pose = SE2(laser_x, laser_y, laser_th);
  r->setIdRobot(id);
  r->setBaseId(10000);
  r->init(0.025, 0.2, 10, 0.15, 2.0, min_inliers);  // cg_mrslam.cpp:69-94 defaults
  r->setInterRobotClosureParams(max_score_mr, min_inliers_mr, window_mr);
  r->currEst = SE2(init_x, init_y, init_th);
  r->odom_prev = SE2(odom_x, odom_y, odom_th);
  r->setInitialData(r->currEst, r->laser(ranges));
  return r;
}

void robot_destroy(void* h) { delete static_cast<Robot*>(h); }

// one keyframe: cg_mrslam.cpp:209-226
void robot_keyframe(void* h, double ox, double oy, double oth, const double* ranges) {
  Robot* r = static_cast<Robot*>(h);
  g2o::trace::select(r->idRobot());
  const SE2 odom(ox, oy, oth);
  r->currEst *= r->odom_prev.inverse() * odom;
  r->odom_prev = odom;
  r->addDataSM(r->currEst, r->laser(ranges));
  r->findConstraints();
  r->findInterRobotConstraints();
  r->optimize(5);
  r->currEst = r->lastVertex()->estimate();
}

int robot_last_vertex(void* h) { return static_cast<Robot*>(h)->lastVertex()->id(); }

// GraphComm::sendToThrd, one period, towards `n_peers` robots in range (graph_comm.cpp:126-150): a
// ComboMessage for all of them if the last vertex changed since the last period, then a
// CondensedGraphMessage per peer that has one due. Datagrams are appended to buf as
// [int32 peer][int32 size][bytes]; returns the bytes used (0 = nothing to send), -1 if buf is short.
int robot_outgoing(void* h, const int* peers, int n_peers, char* buf, int cap) {
  Robot* r = static_cast<Robot*>(h);
  g2o::trace::select(r->idRobot());
  int used = 0;
  static char scratch[MAX_LENGTH_MSG];
  auto append = [&](int peer, RobotMessage* m) -> bool {
    char* end = m->toCharArray(scratch, MAX_LENGTH_MSG);
    const int n = end ? static_cast<int>(end - scratch) : 0;
    if (!n) return true;
    if (used + 8 + n > cap) return false;
    std::memcpy(buf + used, &peer, 4);
    std::memcpy(buf + used + 4, &n, 4);
    std::memcpy(buf + used + 8, scratch, n);
    used += 8 + n;
    return true;
  };
  if (n_peers <= 0) return 0;
  bool ok = true;
  if (r->lastVertex()->id() != r->last_sent_vertex) {
    r->last_sent_vertex = r->lastVertex()->id();
    ComboMessage* cmsg = r->constructComboMessage();
    for (int i = 0; i < n_peers && ok; ++i) ok = append(peers[i], cmsg);
    delete cmsg;
  }
  for (int i = 0; i < n_peers && ok; ++i) {
    CondensedGraphMessage* gmsg = r->constructCondensedGraphMessage(peers[i]);
    if (gmsg) {
      ok = append(peers[i], gmsg);
      delete gmsg;
    }
  }
  return ok ? used : -1;
}

// GraphComm::receiveFromThrd + processQueueThrd for one datagram (graph_comm.cpp:178-211)
int robot_deliver(void* h, const char* data, int n) {
  Robot* r = static_cast<Robot*>(h);
  g2o::trace::select(r->idRobot());
  RobotMessage* msg = r->createMsgfromCharArray(data, n);
  if (!msg) return -1;
  StampedRobotMessage vmsg;
  vmsg.msg = msg;
  vmsg.refVertex = r->lastVertex();
  r->addInterRobotData(vmsg);
  delete msg;
  return 0;
}

// every vertex of this robot's graph: (id, x, y, theta) rows; returns the count (or -needed)
int robot_vertices(void* h, double* out, int cap_rows) {
  Robot* r = static_cast<Robot*>(h);
  const int n = static_cast<int>(r->graph()->vertices().size());
  if (n > cap_rows) return -n;
  int k = 0;
  for (HyperGraph::VertexIDMap::const_iterator it = r->graph()->vertices().begin(); it != r->graph()->vertices().end();
       ++it, ++k) {
    const VertexSE2* v = static_cast<const VertexSE2*>(it->second);
    out[4 * k] = v->id();
    out[4 * k + 1] = v->estimate().translation().x();
    out[4 * k + 2] = v->estimate().translation().y();
    out[4 * k + 3] = v->estimate().rotation().angle();
  }
  return n;
}

// every edge of this robot's graph: (from, to, x, y, theta, information(0,0), level) rows
int robot_edges(void* h, double* out, int cap_rows) {
  Robot* r = static_cast<Robot*>(h);
  const int n = static_cast<int>(r->graph()->edges().size());
  if (n > cap_rows) return -n;
  int k = 0;
  for (HyperGraph::EdgeSet::const_iterator it = r->graph()->edges().begin(); it != r->graph()->edges().end(); ++it, ++k) {
    const EdgeSE2* e = static_cast<const EdgeSE2*>(*it);
    out[7 * k] = e->vertices()[0]->id();
    out[7 * k + 1] = e->vertices()[1]->id();
    out[7 * k + 2] = e->measurement().translation().x();
    out[7 * k + 3] = e->measurement().translation().y();
    out[7 * k + 4] = e->measurement().rotation().angle();
    out[7 * k + 5] = e->information()(0, 0);
    out[7 * k + 6] = e->level();
  }
  return n;
}

// lockstep at the granularity of the solver calls (g2o::trace, include/g2o_compat/g2o_compat.hpp):
// one trace file per robot; the leader records, the follower computes every call itself, notes the
// difference and carries on with the leader's result
int robot_trace_open(void* h, const char* path, int follow) {
  return g2o::trace::open(static_cast<Robot*>(h)->idRobot(), path, follow != 0) ? 0 : -1;
}
// out[0] = index of the call at which the runs parted (-1: they have not), then per kind of call
// (optimize, initial guess, marginals, star measurements, star information): count, worst difference
void robot_trace_stats(void* h, double* out) {
  const g2o::trace::Stream& t = g2o::trace::stats(static_cast<Robot*>(h)->idRobot());
  out[0] = static_cast<double>(t.parted_at);
  for (int k = 1; k < g2o::trace::kKinds; ++k) {
    out[2 * k - 1] = static_cast<double>(t.calls[k]);
    out[2 * k] = t.worst[k];
  }
}
void robot_trace_close(void* h) { g2o::trace::close(static_cast<Robot*>(h)->idRobot()); }

}  // extern "C"
