// TEST INFRASTRUCTURE ONLY. A flat C interface over the message classes of msg_factory.h, compiled
// TWICE from this one source:
//   * against the reference's own msg_factory.h + msg_factory.cpp (verbatim, from /root/reference)
//     -> oracle/_ref/libref_msg.so   (oracle/Makefile, target ref_msg)
//   * against include/cgm/msg_factory.hpp through include/cgm/ref_names/msg_factory.h
//     -> build/libcgm_msg_shim.so    (built by tests/test_msg_wire.py)
// so the tests compare the two implementations byte for byte through the same calls. Only the
// public API of the message classes is used (src/mrslam/msg_factory.h:116-312).
#include <cstring>

#include "msg_factory.h"

namespace {

void fill_vertices(VertexArrayMessage* m, int nv, const int* vid, const double* vest) {
  m->vertexVector.resize(nv);
  for (int i = 0; i < nv; ++i) {
    m->vertexVector[i].id = vid[i];
    for (int k = 0; k < 3; ++k) m->vertexVector[i].estimate[k] = vest[3 * i + k];
  }
}
void fill_laser(RobotLaserMessage* m, int node_id, int nr, const double* readings, const double* laser4) {
  m->nodeId = node_id;
  m->readings.assign(readings, readings + nr);
  m->minangle = laser4[0];
  m->angleincrement = laser4[1];
  m->maxrange = laser4[2];
  m->accuracy = laser4[3];
}
void fill_edges(EdgeArrayMessage* m, int ne, const int* eft, const double* eest, const double* einfo) {
  m->edgeVector.resize(ne);
  for (int i = 0; i < ne; ++i) {
    m->edgeVector[i].idfrom = eft[2 * i];
    m->edgeVector[i].idto = eft[2 * i + 1];
    for (int k = 0; k < 3; ++k) m->edgeVector[i].estimate[k] = eest[3 * i + k];
    for (int k = 0; k < 6; ++k) m->edgeVector[i].information[k] = einfo[6 * i + k];
  }
}
void fill_closures(ClosuresMessage* m, int nc, const int* closures) {
  m->closures.assign(closures, closures + nc);
}

RobotMessage* make(int type) {
  switch (type) {
    case 1: return new VertexArrayMessage();
    case 2: return new RobotLaserMessage();
    case 4: return new ComboMessage();
    case 5: return new EdgeArrayMessage();
    case 6: return new ClosuresMessage();
    case 7: return new CondensedGraphMessage();
    case 8: return new GraphMessage();
  }
  return 0;
}

}  // namespace

extern "C" {

// Serialise one message of `type`; returns the number of bytes written, -1 if it did not fit,
// -2 for an unknown type.
int msg_pack(int type, int robot_id, int nv, const int* vid, const double* vest, int node_id, int nr,
             const double* readings, const double* laser4, int ne, const int* eft, const double* eest,
             const double* einfo, int nc, const int* closures, char* buf, int bsize) {
  RobotMessage* m = make(type);
  if (!m) return -2;
  m->setRobotId(robot_id);
  if (VertexArrayMessage* v = dynamic_cast<VertexArrayMessage*>(m)) fill_vertices(v, nv, vid, vest);
  if (RobotLaserMessage* l = dynamic_cast<RobotLaserMessage*>(m)) fill_laser(l, node_id, nr, readings, laser4);
  if (EdgeArrayMessage* e = dynamic_cast<EdgeArrayMessage*>(m)) fill_edges(e, ne, eft, eest, einfo);
  if (ClosuresMessage* c = dynamic_cast<ClosuresMessage*>(m)) fill_closures(c, nc, closures);
  char* end = m->toCharArray(buf, bsize);
  delete m;
  return end ? static_cast<int>(end - buf) : -1;
}

// Parse one datagram through a MessageFactory with every type registered. counts[6] on entry =
// capacities {nv, nr, ne, nc}, on return = {type, robot id, nv, nr, ne, nc}; returns the bytes
// consumed, -1 if the factory rejects the datagram, -3 if an output array is too small.
int msg_unpack(const char* buf, int size, int* counts, int* vid, double* vest, int* node_id,
               double* readings, double* laser4, int* eft, double* eest, double* einfo, int* closures) {
  MessageFactory f;
  f.registerMessageType<VertexArrayMessage>();
  f.registerMessageType<RobotLaserMessage>();
  f.registerMessageType<ComboMessage>();
  f.registerMessageType<EdgeArrayMessage>();
  f.registerMessageType<ClosuresMessage>();
  f.registerMessageType<CondensedGraphMessage>();
  f.registerMessageType<GraphMessage>();
  const int cap_v = counts[0], cap_r = counts[1], cap_e = counts[2], cap_c = counts[3];
  RobotMessage* m = f.fromCharArray(buf, size);
  if (!m) return -1;
  int out[6] = {m->type(), m->robotId(), 0, 0, 0, 0};
  int rc = size;
  if (VertexArrayMessage* v = dynamic_cast<VertexArrayMessage*>(m)) {
    out[2] = static_cast<int>(v->vertexVector.size());
    if (out[2] > cap_v) rc = -3;
    for (int i = 0; rc > 0 && i < out[2]; ++i) {
      vid[i] = v->vertexVector[i].id;
      for (int k = 0; k < 3; ++k) vest[3 * i + k] = v->vertexVector[i].estimate[k];
    }
  }
  if (RobotLaserMessage* l = dynamic_cast<RobotLaserMessage*>(m)) {
    out[3] = static_cast<int>(l->readings.size());
    if (out[3] > cap_r) rc = -3;
    *node_id = l->nodeId;
    for (int i = 0; rc > 0 && i < out[3]; ++i) readings[i] = l->readings[i];
    laser4[0] = l->minangle;
    laser4[1] = l->angleincrement;
    laser4[2] = l->maxrange;
    laser4[3] = l->accuracy;
  }
  if (EdgeArrayMessage* e = dynamic_cast<EdgeArrayMessage*>(m)) {
    out[4] = static_cast<int>(e->edgeVector.size());
    if (out[4] > cap_e) rc = -3;
    for (int i = 0; rc > 0 && i < out[4]; ++i) {
      eft[2 * i] = e->edgeVector[i].idfrom;
      eft[2 * i + 1] = e->edgeVector[i].idto;
      for (int k = 0; k < 3; ++k) eest[3 * i + k] = e->edgeVector[i].estimate[k];
      for (int k = 0; k < 6; ++k) einfo[6 * i + k] = e->edgeVector[i].information[k];
    }
  }
  if (ClosuresMessage* c = dynamic_cast<ClosuresMessage*>(m)) {
    out[5] = static_cast<int>(c->closures.size());
    if (out[5] > cap_c) rc = -3;
    for (int i = 0; rc > 0 && i < out[5]; ++i) closures[i] = c->closures[i];
  }
  std::memcpy(counts, out, sizeof(out));
  delete m;
  return rc;
}

}  // extern "C"
