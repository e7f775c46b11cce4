// Inter-robot wire format: host-side mirror of the reference's message classes
// (src/mrslam/msg_factory.h:116-312, msg_factory.cpp:30-313) -- SURVEY 8f row 4, first half.
//
// Same class names, members, type ids and byte layout as the reference, so that a front-end built
// against this header exchanges datagrams with a stock cg_mrslam peer:
//   header      int32 type, int32 robotId                       (msg_factory.cpp:30-38)
//   counts      size_t, native width and byte order             (msg_factory.cpp:60-61)
//   ids         int32
//   every double travels as an IEEE float32 (msg_factory.h:78-113): pose estimates, laser ranges
//               and the six upper-triangular information entries lose their low 29 mantissa bits
//               on the wire. That rounding is part of the multi-robot numerics and is reproduced
//               bit for bit (tests/test_msg_wire.py compares against the reference's own
//               msg_factory.cpp compiled verbatim, oracle/_ref/libref_msg.so).
// Composite messages (Combo = vertices + laser, CondensedGraph = edges + closures, Graph = vertices
// + edges + closures) carry ONE header followed by their parts' bodies (msg_factory.cpp:144-158,
// 254-268, 270-291).
//
// Differences, all on error paths: the reference writes past a buffer that is too small for a
// composite message (it re-checks the parts against the whole size, msg_factory.cpp:146-148) and
// never bounds its reads; here toCharArray returns 0 as soon as a field does not fit, and
// fromCharArray returns 0 when a non-zero bsize is exceeded or the type id does not match (the
// reference asserts). MessageFactory::fromCharArray returns 0 instead of asserting on an unknown
// type or a length mismatch, and does not log to cerr. The composite messages' constructors pass
// their robot id on to the (virtual) RobotMessage base; in the reference that argument is lost
// (msg_factory.h:176-180) and every call site sets the id with setRobotId() afterwards.
#ifndef CGM_MSG_FACTORY_HPP
#define CGM_MSG_FACTORY_HPP

#include <cstddef>
#include <cstring>
#include <map>
#include <vector>

#define MAX_LENGTH_MSG 100000

namespace cgm {
namespace wire {

// Cursor over an output buffer; once a field does not fit the cursor is dead and stays dead.
class Writer {
 public:
  Writer(char* at, size_t room) : at_(at), room_(at ? room : 0), ok_(at != 0) {}
  template <typename T>
  void put(const T& v) { raw(&v, sizeof(T)); }
  void put(const double& v) {  // doubles travel as float32
    const float f = static_cast<float>(v);
    raw(&f, sizeof(float));
  }
  char* end() const { return ok_ ? at_ : 0; }
  size_t room() const { return ok_ ? room_ : 0; }

 private:
  void raw(const void* src, size_t n) {
    if (!ok_ || n > room_) {
      ok_ = false;
      return;
    }
    std::memcpy(at_, src, n);
    at_ += n;
    room_ -= n;
  }
  char* at_;
  size_t room_;
  bool ok_;
};

// Cursor over an input buffer; limit 0 = unbounded (the reference's behaviour).
class Reader {
 public:
  Reader(const char* at, size_t limit) : at_(at), left_(limit), bounded_(limit > 0), ok_(at != 0) {}
  template <typename T>
  void get(T& v) { raw(&v, sizeof(T)); }
  void get(double& v) {
    float f = 0.f;
    raw(&f, sizeof(float));
    v = f;
  }
  const char* end() const { return ok_ ? at_ : 0; }
  size_t left() const { return bounded_ ? left_ : 0; }
  bool ok() const { return ok_; }
  void fail() { ok_ = false; }
  // a count that cannot possibly be backed by the remaining bytes (bounded reads only)
  bool plausible(size_t count, size_t bytes_each) const {
    return !bounded_ || count <= left_ / (bytes_each ? bytes_each : 1);
  }

 private:
  void raw(void* dst, size_t n) {
    if (!ok_ || (bounded_ && n > left_)) {
      ok_ = false;
      return;
    }
    std::memcpy(dst, at_, n);
    at_ += n;
    if (bounded_) left_ -= n;
  }
  const char* at_;
  size_t left_;
  bool bounded_, ok_;
};

}  // namespace wire
}  // namespace cgm

struct RobotMessage {
  RobotMessage(int robotId_ = -1) : _robotId(robotId_) {}
  virtual ~RobotMessage() {}

  // serialises the message into the buffer; returns one past the last byte written, or 0
  virtual char* toCharArray(char* buffer, size_t bsize, bool skipHeader = false) const {
    cgm::wire::Writer w(buffer, bsize);
    if (!skipHeader) putHeader(w);
    putBody(w);
    return w.end();
  }
  // fills in the data fields from the buffer; returns one past the last byte read, or 0
  virtual const char* fromCharArray(const char* buffer, size_t bsize = 0, bool skipHeader = false) {
    cgm::wire::Reader r(buffer, bsize);
    if (!skipHeader) getHeader(r);
    if (r.ok()) getBody(r);
    return r.end();
  }

  void setRobotId(int rid) { _robotId = rid; }
  int robotId() const { return _robotId; }
  static int _type() { return 0; }
  virtual int type() const = 0;

 protected:
  void putHeader(cgm::wire::Writer& w) const {
    w.put(type());
    w.put(_robotId);
  }
  void getHeader(cgm::wire::Reader& r) {
    int t = -1;
    r.get(t);
    r.get(_robotId);
    if (t != type()) r.fail();
  }
  // the parts a concrete message is made of, in wire order
  virtual void putBody(cgm::wire::Writer&) const {}
  virtual void getBody(cgm::wire::Reader&) {}

  int _robotId;
};

struct VertexArrayMessage : public virtual RobotMessage {
  struct VSE2Data {
    int id;
    double estimate[3];
  };
  VertexArrayMessage(int robotId_ = -1) : RobotMessage(robotId_) {}

  std::vector<VSE2Data> vertexVector;
  static int _type() { return 1; }
  virtual int type() const { return _type(); }

 protected:
  void putVertices(cgm::wire::Writer& w) const {
    w.put(static_cast<size_t>(vertexVector.size()));
    for (size_t i = 0; i < vertexVector.size(); ++i) {
      w.put(vertexVector[i].id);
      for (int k = 0; k < 3; ++k) w.put(vertexVector[i].estimate[k]);
    }
  }
  void getVertices(cgm::wire::Reader& r) {
    size_t n = 0;
    r.get(n);
    if (!r.ok() || !r.plausible(n, sizeof(int) + 3 * sizeof(float))) return r.fail();
    vertexVector.resize(n);
    for (size_t i = 0; i < n; ++i) {
      r.get(vertexVector[i].id);
      for (int k = 0; k < 3; ++k) r.get(vertexVector[i].estimate[k]);
    }
  }
  virtual void putBody(cgm::wire::Writer& w) const { putVertices(w); }
  virtual void getBody(cgm::wire::Reader& r) { getVertices(r); }
};

struct RobotLaserMessage : public virtual RobotMessage {
  RobotLaserMessage(int robotId_ = -1)
      : RobotMessage(robotId_), nodeId(-1), minangle(0), angleincrement(0), maxrange(0), accuracy(0) {}

  int nodeId;
  std::vector<double> readings;
  // laser params
  double minangle;
  double angleincrement;
  double maxrange;
  double accuracy;

  static int _type() { return 2; }
  virtual int type() const { return _type(); }

 protected:
  void putLaser(cgm::wire::Writer& w) const {
    w.put(nodeId);
    w.put(static_cast<size_t>(readings.size()));
    for (size_t i = 0; i < readings.size(); ++i) w.put(readings[i]);
    w.put(minangle);
    w.put(angleincrement);
    w.put(maxrange);
    w.put(accuracy);
  }
  void getLaser(cgm::wire::Reader& r) {
    size_t n = 0;
    r.get(nodeId);
    r.get(n);
    if (!r.ok() || !r.plausible(n, sizeof(float))) return r.fail();
    readings.resize(n);
    for (size_t i = 0; i < n; ++i) r.get(readings[i]);
    r.get(minangle);
    r.get(angleincrement);
    r.get(maxrange);
    r.get(accuracy);
  }
  virtual void putBody(cgm::wire::Writer& w) const { putLaser(w); }
  virtual void getBody(cgm::wire::Reader& r) { getLaser(r); }
};

struct ComboMessage : public VertexArrayMessage, RobotLaserMessage {
  ComboMessage(int robotId_ = -1)
      : RobotMessage(robotId_), VertexArrayMessage(robotId_), RobotLaserMessage(robotId_) {}
  static int _type() { return 4; }
  virtual int type() const { return _type(); }

 protected:
  virtual void putBody(cgm::wire::Writer& w) const {
    putVertices(w);
    putLaser(w);
  }
  virtual void getBody(cgm::wire::Reader& r) {
    getVertices(r);
    if (r.ok()) getLaser(r);
  }
};

struct EdgeArrayMessage : public virtual RobotMessage {
  struct ESE2Data {
    int idfrom;
    int idto;
    double estimate[3];
    double information[6];
  };
  EdgeArrayMessage(int robotId_ = -1) : RobotMessage(robotId_) {}

  std::vector<ESE2Data> edgeVector;
  static int _type() { return 5; }
  virtual int type() const { return _type(); }

 protected:
  void putEdges(cgm::wire::Writer& w) const {
    w.put(static_cast<size_t>(edgeVector.size()));
    for (size_t i = 0; i < edgeVector.size(); ++i) {
      const ESE2Data& e = edgeVector[i];
      w.put(e.idfrom);
      w.put(e.idto);
      for (int k = 0; k < 3; ++k) w.put(e.estimate[k]);
      for (int k = 0; k < 6; ++k) w.put(e.information[k]);
    }
  }
  void getEdges(cgm::wire::Reader& r) {
    size_t n = 0;
    r.get(n);
    if (!r.ok() || !r.plausible(n, 2 * sizeof(int) + 9 * sizeof(float))) return r.fail();
    edgeVector.resize(n);
    for (size_t i = 0; i < n; ++i) {
      ESE2Data& e = edgeVector[i];
      r.get(e.idfrom);
      r.get(e.idto);
      for (int k = 0; k < 3; ++k) r.get(e.estimate[k]);
      for (int k = 0; k < 6; ++k) r.get(e.information[k]);
    }
  }
  virtual void putBody(cgm::wire::Writer& w) const { putEdges(w); }
  virtual void getBody(cgm::wire::Reader& r) { getEdges(r); }
};

struct ClosuresMessage : public virtual RobotMessage {
  ClosuresMessage(int robotId_ = -1) : RobotMessage(robotId_) {}

  std::vector<int> closures;
  static int _type() { return 6; }
  virtual int type() const { return _type(); }

 protected:
  void putClosures(cgm::wire::Writer& w) const {
    w.put(static_cast<size_t>(closures.size()));
    for (size_t i = 0; i < closures.size(); ++i) w.put(closures[i]);
  }
  void getClosures(cgm::wire::Reader& r) {
    size_t n = 0;
    r.get(n);
    if (!r.ok() || !r.plausible(n, sizeof(int))) return r.fail();
    closures.resize(n);
    for (size_t i = 0; i < n; ++i) r.get(closures[i]);
  }
  virtual void putBody(cgm::wire::Writer& w) const { putClosures(w); }
  virtual void getBody(cgm::wire::Reader& r) { getClosures(r); }
};

struct CondensedGraphMessage : public EdgeArrayMessage, ClosuresMessage {
  CondensedGraphMessage(int robotId_ = -1)
      : RobotMessage(robotId_), EdgeArrayMessage(robotId_), ClosuresMessage(robotId_) {}
  static int _type() { return 7; }
  virtual int type() const { return _type(); }

 protected:
  virtual void putBody(cgm::wire::Writer& w) const {
    putEdges(w);
    putClosures(w);
  }
  virtual void getBody(cgm::wire::Reader& r) {
    getEdges(r);
    if (r.ok()) getClosures(r);
  }
};

struct GraphMessage : public VertexArrayMessage, EdgeArrayMessage, ClosuresMessage {
  GraphMessage(int robotId_ = -1)
      : RobotMessage(robotId_), VertexArrayMessage(robotId_), EdgeArrayMessage(robotId_),
        ClosuresMessage(robotId_) {}
  static int _type() { return 8; }
  virtual int type() const { return _type(); }

 protected:
  virtual void putBody(cgm::wire::Writer& w) const {
    putVertices(w);
    putEdges(w);
    putClosures(w);
  }
  virtual void getBody(cgm::wire::Reader& r) {
    getVertices(r);
    if (r.ok()) getEdges(r);
    if (r.ok()) getClosures(r);
  }
};

// Registry type id -> constructor (msg_factory.h:269-311); owns its creators.
struct MessageFactory {
  size_t size() const { return _makers.size(); }

  template <typename T>
  void registerMessageType() {
    // the reference asserts on a duplicate registration; registering twice is harmless here
    _makers[T::_type()] = &MessageFactory::make<T>;
  }
  RobotMessage* constructMessage(int t) {
    std::map<int, Creator>::const_iterator it = _makers.find(t);
    return it == _makers.end() ? 0 : (it->second)();
  }
  // a new message of the type found in the datagram's header, filled from it; 0 if the type is
  // unknown or the datagram is not exactly one message long. The caller owns the result.
  RobotMessage* fromCharArray(const char* buf, size_t size) {
    if (!buf || size <= sizeof(int)) return 0;
    int t = -1;
    std::memcpy(&t, buf, sizeof(int));
    RobotMessage* msg = constructMessage(t);
    if (!msg) return 0;
    const char* end = msg->fromCharArray(buf, size);
    if (!end || static_cast<size_t>(end - buf) != size) {
      delete msg;
      return 0;
    }
    return msg;
  }

 private:
  typedef RobotMessage* (*Creator)();
  template <typename T>
  static RobotMessage* make() { return new T(); }
  std::map<int, Creator> _makers;
};

#endif
