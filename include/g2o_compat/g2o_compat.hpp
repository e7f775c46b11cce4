// g2o-compatible host layer: exactly the subset of g2o's public classes that the reference
// touches (enumerated in SURVEY.md section 8b), so that src/slam/*.cpp, src/mrslam/*.cpp and
// src/matcher/scan_matcher.cpp-style code compiles against it. The graph container and the SE(2)
// value types are plain host C++; optimize() / computeMarginals() / computeInitialGuess() /
// EdgeLabeler::labelEdges() marshal the active set into flat arrays and call the CUDA solver
// through include/pgo_solver.h. There is no CPU solver here.
//
// g2o is not in the container; behaviour follows SURVEY.md appendix C ("recalled") and the
// reference's call sites cited per member. Deliberate, documented deviation:
// VertexSet / EdgeSet are ordered by id / insertion index instead of by pointer value, which
// makes g2o's pointer-order-dependent choices deterministic.
#ifndef G2O_COMPAT_HPP
#define G2O_COMPAT_HPP

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <queue>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "../cgm/device.hpp"
#include "../cgm/eigen_lite.hpp"
#include "../pgo_solver.h"

namespace g2o {

// ---- call trace (test instrumentation, off unless a driver opens it) ------------------------------
// The reference's front-end takes discrete decisions from solver outputs (closest vertex, covariance
// gate, closure votes); two correct solvers that agree to 1e-13 can still break an exact tie
// differently, after which two free runs are two different SLAM sessions. To compare a whole bag the
// tests run two builds of the same sources in LOCKSTEP at the granularity of the solver calls: the
// leader (this header over the CUDA library) records the result of every optimize /
// computeInitialGuess / computeMarginals / labelEdges; the follower (the same header over the CPU
// oracle solver) computes its own result from the same inputs, the trace notes the difference, and
// the follower carries on with the leader's numbers. Records are keyed by vertex id, so that
// pointer-ordered containers on either side do not matter.
namespace trace {
enum Kind { kOptimize = 1, kInitialGuess = 2, kMarginals = 3, kStarMeas = 4, kStarInfo = 5, kKinds = 6 };
struct Stream {
  FILE* f = nullptr;
  int mode = 0;                   // 0 off, 1 record, 2 follow
  long long calls[kKinds] = {0, 0, 0, 0, 0, 0};
  double worst[kKinds] = {0, 0, 0, 0, 0, 0};   // poses, measurements: absolute (angles wrapped);
                                               // covariance / information blocks: relative to the block's largest entry
  long long parted_at = -1;       // index of the first call whose keys differed (the runs have parted)
  long long n_calls = 0;
};
inline Stream* streams() {
  static Stream s[16];
  return s;
}
inline int& current() {
  static int c = 0;
  return c;
}
inline bool open(int stream, const char* path, bool follow) {
  Stream& t = streams()[stream & 15];
  if (t.f) std::fclose(t.f);
  t = Stream();
  t.f = std::fopen(path, follow ? "rb" : "wb");
  t.mode = t.f ? (follow ? 2 : 1) : 0;
  return t.f != nullptr;
}
inline void select(int stream) { current() = stream & 15; }
inline void close(int stream) {
  Stream& t = streams()[stream & 15];
  if (t.f) std::fclose(t.f);
  t.f = nullptr;
  t.mode = 0;
}
inline const Stream& stats(int stream) { return streams()[stream & 15]; }
// wall-clock milliseconds spent in the solver library, by entry point (always on: two clock reads per
// call): [0] pgo_set_graph, [1] pgo_upload, [2] pgo_iterate, [3] pgo_marginals, [4] pgo_initial_guess
// (+ pose read-back), [5] pgo_label_star_edges; [6 + k]: the number of calls of entry k
inline double* wall_ms() {
  static double t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  return t;
}
struct Clock {
  int k;
  std::chrono::steady_clock::time_point t0;
  explicit Clock(int kind) : k(kind), t0(std::chrono::steady_clock::now()) {}
  ~Clock() {
    wall_ms()[k] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    wall_ms()[6 + k] += 1.0;
  }
};

// One solver result: `keys.size()` rows of `width` doubles. angle_col: column holding an angle (or
// -1); relative: compare relative to the row's largest magnitude.
inline void exchange(int kind, const std::vector<long long>& keys, int width, std::vector<double>& values,
                     int angle_col, bool relative) {
  Stream& t = streams()[current()];
  if (!t.mode) return;
  const long long n = static_cast<long long>(keys.size());
  ++t.n_calls;
  if (t.mode == 1) {
    const long long head[3] = {kind, n, width};
    std::fwrite(head, sizeof head, 1, t.f);
    if (n) {
      std::fwrite(keys.data(), sizeof(long long), n, t.f);
      std::fwrite(values.data(), sizeof(double), n * width, t.f);
    }
    ++t.calls[kind];
    return;
  }
  long long head[3] = {0, 0, 0};
  std::vector<long long> lk;
  std::vector<double> lv;
  bool ok = std::fread(head, sizeof head, 1, t.f) == 1 && head[0] == kind && head[1] == n && head[2] == width;
  if (ok && n) {
    lk.resize(n);
    lv.resize(n * width);
    ok = std::fread(lk.data(), sizeof(long long), n, t.f) == static_cast<size_t>(n) &&
         std::fread(lv.data(), sizeof(double), n * width, t.f) == static_cast<size_t>(n * width);
  }
  std::map<long long, long long> row_of;
  if (ok) {
    for (long long i = 0; i < n; ++i) row_of[lk[i]] = i;
    std::set<long long> own(keys.begin(), keys.end());
    for (long long i = 0; i < n && ok; ++i) ok = row_of.count(keys[i]) != 0;
    ok = ok && own.size() == row_of.size();   // the same set of keys on both sides
  }
  if (!ok) {  // not the call the leader made here: the runs have parted, stop following
    t.parted_at = t.n_calls - 1;
    t.mode = 0;
    return;
  }
  const double two_pi = 6.283185307179586476925286766559;
  for (long long i = 0; i < n; ++i) {
    const double* theirs = lv.data() + row_of[keys[i]] * width;
    double* mine = values.data() + i * width;
    double scale = 1.0;
    if (relative) {
      scale = 0.0;
      for (int c = 0; c < width; ++c) scale = std::max(scale, std::fabs(theirs[c]));
      if (!(scale > 0.0)) scale = 1.0;
    }
    for (int c = 0; c < width; ++c) {
      double d = mine[c] - theirs[c];
      if (c == angle_col) d -= two_pi * std::floor(d / two_pi + 0.5);
      d = std::fabs(d) / scale;
      if (!(d <= t.worst[kind])) t.worst[kind] = d;   // also catches NaN
      mine[c] = theirs[c];
    }
  }
  ++t.calls[kind];
}
}  // namespace trace

// ---- C2: normalize_theta ------------------------------------------------------------------------
inline double normalize_theta(double theta) {
  if (theta >= -M_PI && theta < M_PI) return theta;
  const double two_pi = 2.0 * M_PI;
  double w = theta - two_pi * std::floor(theta / two_pi);
  if (w >= M_PI) w -= two_pi;
  return w;
}

// ---- C1: SE2 ------------------------------------------------------------------------------------
class SE2 {
 public:
  SE2() : t_(0.0, 0.0), r_(0.0) {}
  SE2(double x, double y, double theta) : t_(x, y), r_(theta) {}
  const Eigen::Vector2d& translation() const { return t_; }
  void setTranslation(const Eigen::Vector2d& t) { t_ = t; }
  const Eigen::Rotation2Dd& rotation() const { return r_; }
  void setRotation(const Eigen::Rotation2Dd& r) { r_ = r; }
  SE2 operator*(const SE2& o) const {
    SE2 r;
    r.t_ = t_ + (r_ * o.t_);
    r.r_ = Eigen::Rotation2Dd(normalize_theta(r_.angle() + o.r_.angle()));
    return r;
  }
  SE2& operator*=(const SE2& o) {
    *this = *this * o;
    return *this;
  }
  Eigen::Vector2d operator*(const Eigen::Vector2d& v) const { return t_ + (r_ * v); }
  SE2 inverse() const {
    SE2 r;
    r.r_ = Eigen::Rotation2Dd(normalize_theta(-r_.angle()));
    r.t_ = r.r_ * (t_ * -1.0);
    return r;
  }
  Eigen::Vector3d toVector() const { return Eigen::Vector3d(t_.x(), t_.y(), r_.angle()); }
  void fromVector(const Eigen::Vector3d& v) { *this = SE2(v[0], v[1], v[2]); }
  double operator[](int i) const { return i < 2 ? t_[i] : r_.angle(); }

 private:
  Eigen::Vector2d t_;
  Eigen::Rotation2Dd r_;
};

// ---- graph container ------------------------------------------------------------------------------
class SparseOptimizer;

struct HyperGraph {
  class Edge;
  // user data attached to a vertex: a singly linked list (g2o HyperGraph::Data / OptimizableGraph::
  // Data; walked by GraphSLAM::findLaserData, graph_slam.cpp:271-284)
  class Data {
   public:
    Data() : next_(nullptr) {}
    virtual ~Data() { delete next_; }
    Data* next() const { return next_; }
    void setNext(Data* n) { next_ = n; }

   private:
    Data* next_;
  };
  class Vertex {
   public:
    Vertex() : id_(-1) {}
    virtual ~Vertex() {}
    int id() const { return id_; }
    virtual void setId(int id) { id_ = id; }
    struct EdgeLess {
      bool operator()(const Edge* a, const Edge* b) const;
    };
    typedef std::set<Edge*, EdgeLess> EdgeSetT;
    const EdgeSetT& edges() const { return edges_; }
    EdgeSetT& edges() { return edges_; }

   protected:
    int id_;
    EdgeSetT edges_;
  };
  struct VertexLess {
    bool operator()(const Vertex* a, const Vertex* b) const {
      return a->id() != b->id() ? a->id() < b->id() : a < b;
    }
  };
  class Edge {
   public:
    // serial: creation index, immutable -- the key of every EdgeSet (g2o orders its sets by pointer
    // value; an index that never changes gives the same container semantics, deterministically).
    // internalId: insertion index into a graph (g2o assigns it in addEdge; active edges are
    // processed in this order, C5)
    Edge() : id_(-1), serial_(nextSerial()++), internal_id_(-1) { vertices_.resize(2, nullptr); }
    virtual ~Edge() {}
    std::vector<Vertex*>& vertices() { return vertices_; }
    const std::vector<Vertex*>& vertices() const { return vertices_; }
    Vertex* vertex(size_t i) const { return vertices_[i]; }
    void setVertex(size_t i, Vertex* v) { vertices_[i] = v; }
    int id() const { return id_; }
    void setId(int id) { id_ = id; }
    long long serial() const { return serial_; }
    long long internalId() const { return internal_id_; }
    void setInternalId(long long s) { internal_id_ = s; }

   protected:
    static long long& nextSerial() {
      static long long n = 0;
      return n;
    }
    std::vector<Vertex*> vertices_;
    int id_;
    long long serial_, internal_id_;
  };
  // ordered by vertex id; sets of VertexSets (graph_slam.cpp:404, vertices_finder.cpp:60-112) compare
  // by ids too, so their iteration order does not depend on heap addresses (with g2o it does)
  struct VertexSet : public std::set<Vertex*, VertexLess> {
    friend bool operator<(const VertexSet& a, const VertexSet& b) {
      const_iterator i = a.begin(), j = b.begin();
      for (; i != a.end() && j != b.end(); ++i, ++j) {
        if ((*i)->id() != (*j)->id()) return (*i)->id() < (*j)->id();
        if (*i != *j) return *i < *j;
      }
      return i == a.end() && j != b.end();
    }
  };
  typedef Vertex::EdgeSetT EdgeSet;
  typedef std::map<int, Vertex*> VertexIDMap;

  virtual ~HyperGraph() {}
  Vertex* vertex(int id) const {
    VertexIDMap::const_iterator it = vertices_.find(id);
    return it == vertices_.end() ? nullptr : it->second;
  }
  const VertexIDMap& vertices() const { return vertices_; }
  VertexIDMap& vertices() { return vertices_; }
  const EdgeSet& edges() const { return edges_; }
  EdgeSet& edges() { return edges_; }

 protected:
  VertexIDMap vertices_;
  EdgeSet edges_;
};

inline bool HyperGraph::Vertex::EdgeLess::operator()(const Edge* a, const Edge* b) const {
  return a->serial() != b->serial() ? a->serial() < b->serial() : a < b;
}

struct OptimizableGraph : public HyperGraph {
  typedef HyperGraph::Data Data;
  class Vertex : public HyperGraph::Vertex {
   public:
    Vertex() : fixed_(false), hessian_index_(-1), user_data_(nullptr) {}
    virtual ~Vertex() { delete user_data_; }  // deletes the whole chain
    bool fixed() const { return fixed_; }
    void setFixed(bool f) { fixed_ = f; }
    int hessianIndex() const { return hessian_index_; }
    void setHessianIndex(int h) { hessian_index_ = h; }
    HyperGraph::Data* userData() const { return user_data_; }
    void setUserData(HyperGraph::Data* d) { user_data_ = d; }  // the vertex owns its user data
    void addUserData(HyperGraph::Data* d) {  // g2o prepends: the newest datum is userData()
      if (d) {
        d->setNext(user_data_);
        user_data_ = d;
      }
    }
    virtual void push() = 0;
    virtual void pop() = 0;

   protected:
    bool fixed_;
    int hessian_index_;
    HyperGraph::Data* user_data_;
  };
  class Edge : public HyperGraph::Edge {
   public:
    Edge() : level_(0) {}
    int level() const { return level_; }
    void setLevel(int l) { level_ = l; }
    virtual void computeError() = 0;
    virtual double chi2() const = 0;

   protected:
    int level_;
  };
  Vertex* vertex(int id) const { return static_cast<Vertex*>(HyperGraph::vertex(id)); }
};

// ---- C13: laser data ------------------------------------------------------------------------------
struct LaserParameters {
  LaserParameters(int type_, int beams, double firstBeamAngle_, double angularStep_,
                  double maxRange_, double accuracy_, int remissionMode_)
      : laserPose(), type(type_), firstBeamAngle(firstBeamAngle_), fov(angularStep_ * beams),
        angularStep(angularStep_), accuracy(accuracy_), remissionMode(remissionMode_),
        maxRange(maxRange_) {}
  LaserParameters() : LaserParameters(0, 0, 0.0, 0.0, 0.0, 0.0, 0) {}
  SE2 laserPose;
  int type;
  double firstBeamAngle, fov, angularStep, accuracy;
  int remissionMode;
  double maxRange;
};

class RawLaser : public HyperGraph::Data {
 public:
  typedef std::vector<Eigen::Vector2d, Eigen::aligned_allocator<Eigen::Vector2d> > Point2DVector;
  virtual ~RawLaser() {}
  // beams with r < maxRange, in the LASER frame (the reference applies laserPose itself,
  // scan_matcher.cpp:98-106)
  Point2DVector cartesian() const {
    Point2DVector pts;
    for (size_t i = 0; i < ranges_.size(); ++i) {
      const double r = ranges_[i];
      if (r < params_.maxRange) {
        const double a = params_.firstBeamAngle + i * params_.angularStep;
        pts.push_back(Eigen::Vector2d(std::cos(a) * r, std::sin(a) * r));
      }
    }
    return pts;
  }
  const std::vector<double>& ranges() const { return ranges_; }
  void setRanges(const std::vector<double>& r) { ranges_ = r; }
  const std::vector<double>& remissions() const { return remissions_; }
  void setRemissions(const std::vector<double>& r) { remissions_ = r; }
  const LaserParameters& laserParams() const { return params_; }
  void setLaserParams(const LaserParameters& p) { params_ = p; }

 protected:
  std::vector<double> ranges_, remissions_;
  LaserParameters params_;
};

class RobotLaser : public RawLaser {
 public:
  RobotLaser() : timestamp_(0.0), logger_timestamp_(0.0) {}
  const SE2& odomPose() const { return odom_; }
  void setOdomPose(const SE2& p) { odom_ = p; }
  double timestamp() const { return timestamp_; }
  void setTimestamp(double t) { timestamp_ = t; }
  double loggerTimestamp() const { return logger_timestamp_; }
  void setLoggerTimestamp(double t) { logger_timestamp_ = t; }
  const std::string& hostname() const { return hostname_; }
  void setHostname(const std::string& h) { hostname_ = h; }

 private:
  SE2 odom_;
  double timestamp_, logger_timestamp_;
  std::string hostname_;
};

class VertexSE2 : public OptimizableGraph::Vertex {
 public:
  const SE2& estimate() const { return estimate_; }
  void setEstimate(const SE2& e) { estimate_ = e; }
  void push() override { stack_.push_back(estimate_); }
  void pop() override {
    if (!stack_.empty()) {
      estimate_ = stack_.back();
      stack_.pop_back();
    }
  }
  void oplus(const double* d) {  // C3: additive, world frame
    estimate_ = SE2(estimate_.translation().x() + d[0], estimate_.translation().y() + d[1],
                    normalize_theta(estimate_.rotation().angle() + d[2]));
  }

 private:
  SE2 estimate_;
  std::vector<SE2> stack_;
};

class EdgeSE2 : public OptimizableGraph::Edge {
 public:
  EdgeSE2() : information_(Eigen::Matrix3d::Identity()) {}
  const SE2& measurement() const { return measurement_; }
  void setMeasurement(const SE2& m) {
    measurement_ = m;
    inverse_measurement_ = m.inverse();
  }
  const Eigen::Matrix3d& information() const { return information_; }
  Eigen::Matrix3d& information() { return information_; }
  void setInformation(const Eigen::Matrix3d& i) { information_ = i; }
  void computeError() override {  // C4
    const VertexSE2* v1 = static_cast<const VertexSE2*>(vertices_[0]);
    const VertexSE2* v2 = static_cast<const VertexSE2*>(vertices_[1]);
    error_ = (inverse_measurement_ * (v1->estimate().inverse() * v2->estimate())).toVector();
  }
  const Eigen::Vector3d& error() const { return error_; }
  double chi2() const override {
    const Eigen::Vector3d we = information_ * error_;
    return error_[0] * we[0] + error_[1] * we[1] + error_[2] * we[2];
  }
  void setMeasurementFromState() {
    const VertexSE2* v1 = static_cast<const VertexSE2*>(vertices_[0]);
    const VertexSE2* v2 = static_cast<const VertexSE2*>(vertices_[1]);
    setMeasurement(v1->estimate().inverse() * v2->estimate());
  }

 private:
  SE2 measurement_, inverse_measurement_;
  Eigen::Matrix3d information_;
  Eigen::Vector3d error_;
};

// SparseBlockMatrix<MatrixXd>: what computeMarginals fills (graph_manipulator.cpp:134-155).
template <typename MatrixType>
class SparseBlockMatrix {
 public:
  MatrixType* block(int r, int c, bool alloc = false) {
    typename std::map<std::pair<int, int>, MatrixType>::iterator it = blocks_.find(std::make_pair(r, c));
    if (it != blocks_.end()) return &it->second;
    if (!alloc) return nullptr;
    return &(blocks_[std::make_pair(r, c)] = MatrixType(3, 3));
  }
  void clear() { blocks_.clear(); }

 private:
  std::map<std::pair<int, int>, MatrixType> blocks_;
};

// ---- construction idiom kept compiling (graph_slam.cpp:44-55); the objects carry no behaviour ---
template <int P, int L>
struct BlockSolverTraits {
  typedef Eigen::MatrixXd PoseMatrixType;
};
template <typename MatrixType>
struct LinearSolverCSparse {
  void setBlockOrdering(bool) {}
};
struct Solver {
  virtual ~Solver() {}
};
template <typename Traits>
struct BlockSolver : public Solver {
  typedef typename Traits::PoseMatrixType PoseMatrixType;
  typedef LinearSolverCSparse<PoseMatrixType> LinearSolverType;
  explicit BlockSolver(std::unique_ptr<LinearSolverType>) {}
  explicit BlockSolver(LinearSolverType* p) { delete p; }
};
struct OptimizationAlgorithm {
  virtual ~OptimizationAlgorithm() {}
};
struct OptimizationAlgorithmGaussNewton : public OptimizationAlgorithm {
  explicit OptimizationAlgorithmGaussNewton(std::unique_ptr<Solver>) {}
  explicit OptimizationAlgorithmGaussNewton(Solver* p) { delete p; }
};

// ---- HyperDijkstra --------------------------------------------------------------------------------
// g2o::HyperDijkstra as the reference uses it (vertices_finder.cpp:37-41,46-50,87-93):
// uniform-cost search from one vertex over ALL edges incident to a vertex (any level, both
// directions), edge cost from a user functor (max() = impassable), a vertex is reached when the new
// distance is below its old one and below maxDistance; visited() = every vertex reached, the source
// included. g2o (recalled) compares with a conditioner, d_new + 1e-3 < d_old, kept as its default
// parameter. Ties in the queue are broken by vertex id, so the result is deterministic (g2o's own
// order depends on pointer values; the result SET does not).
class HyperDijkstra {
 public:
  struct CostFunction {
    virtual double operator()(HyperGraph::Edge* e, HyperGraph::Vertex* from, HyperGraph::Vertex* to) = 0;
    virtual ~CostFunction() {}
  };
  explicit HyperDijkstra(HyperGraph* g) : graph_(g) {}
  HyperGraph::VertexSet& visited() { return visited_; }
  double distance(HyperGraph::Vertex* v) const {
    std::map<HyperGraph::Vertex*, double>::const_iterator it = dist_.find(v);
    return it == dist_.end() ? std::numeric_limits<double>::max() : it->second;
  }
  void shortestPaths(HyperGraph::Vertex* source, CostFunction* cost,
                     double maxDistance = std::numeric_limits<double>::max(),
                     double comparisonConditioner = 1e-3, bool directed = false,
                     double maxEdgeCost = std::numeric_limits<double>::max()) {
    visited_.clear();
    dist_.clear();
    typedef std::pair<double, std::pair<int, HyperGraph::Vertex*> > Entry;  // (distance, id, vertex)
    std::priority_queue<Entry, std::vector<Entry>, std::greater<Entry> > queue;
    dist_[source] = 0.0;
    visited_.insert(source);
    queue.push(Entry(0.0, std::make_pair(source->id(), source)));
    while (!queue.empty()) {
      const Entry top = queue.top();
      queue.pop();
      HyperGraph::Vertex* u = top.second.second;
      if (top.first > dist_[u]) continue;  // stale entry
      for (HyperGraph::EdgeSet::const_iterator it = u->edges().begin(); it != u->edges().end(); ++it) {
        HyperGraph::Edge* e = *it;
        if (directed && e->vertex(0) != u) continue;
        for (size_t k = 0; k < e->vertices().size(); ++k) {
          HyperGraph::Vertex* z = e->vertex(k);
          if (!z || z == u) continue;
          const double c = (*cost)(e, u, z);
          if (c == std::numeric_limits<double>::max() || c > maxEdgeCost) continue;
          const double nd = top.first + c;
          if (nd + comparisonConditioner < distance(z) && nd < maxDistance) {
            dist_[z] = nd;
            visited_.insert(z);
            queue.push(Entry(nd, std::make_pair(z->id(), z)));
          }
        }
      }
    }
    (void)graph_;
  }

 private:
  HyperGraph* graph_;
  HyperGraph::VertexSet visited_;
  std::map<HyperGraph::Vertex*, double> dist_;
};

// ---- SparseOptimizer ------------------------------------------------------------------------------
class SparseOptimizer : public OptimizableGraph {
 public:
  typedef std::vector<OptimizableGraph::Vertex*> VertexContainer;
  typedef std::vector<OptimizableGraph::Edge*> EdgeContainer;

  explicit SparseOptimizer(int device = -1)
      : solver_(nullptr), algorithm_(nullptr), verbose_(false), next_serial_(0),
        device_(device >= 0 ? device : cgm::default_device()), structure_dirty_(true), use_clock_(0) {
    for (int k = 0; k < kSolverSlots; ++k) {
      slots_[k] = nullptr;
      slot_key_[k] = 0;
      slot_used_[k] = 0;
    }
  }
  ~SparseOptimizer() {
    for (int k = 0; k < kSolverSlots; ++k)
      if (slots_[k]) pgo_destroy(slots_[k]);
    delete algorithm_;
    for (HyperGraph::EdgeSet::iterator it = edges_.begin(); it != edges_.end(); ++it) delete *it;
    for (VertexIDMap::iterator it = vertices_.begin(); it != vertices_.end(); ++it) delete it->second;
  }
  SparseOptimizer(const SparseOptimizer&) = delete;
  SparseOptimizer& operator=(const SparseOptimizer&) = delete;

  void setAlgorithm(OptimizationAlgorithm* a) {
    delete algorithm_;
    algorithm_ = a;
  }
  void setVerbose(bool v) { verbose_ = v; }

  // the graph owns vertices and edges after add* (g2o convention)
  bool addVertex(OptimizableGraph::Vertex* v) {
    if (vertices_.count(v->id())) return false;
    vertices_[v->id()] = v;
    structure_dirty_ = true;
    return true;
  }
  bool addEdge(OptimizableGraph::Edge* e) {
    if (!e->vertex(0) || !e->vertex(1)) return false;
    e->setInternalId(next_serial_++);
    edges_.insert(e);
    e->vertex(0)->edges().insert(e);
    e->vertex(1)->edges().insert(e);
    structure_dirty_ = true;
    return true;
  }
  bool removeEdge(HyperGraph::Edge* e) {  // deletes the edge (g2o convention)
    if (!edges_.erase(e)) return false;
    for (size_t k = 0; k < e->vertices().size(); ++k)
      if (e->vertex(k)) e->vertex(k)->edges().erase(e);
    delete e;
    structure_dirty_ = true;
    return true;
  }
  const VertexContainer& activeVertices() const { return active_vertices_; }
  const EdgeContainer& activeEdges() const { return active_edges_; }

  // C6: active edges = level-`level` edges; active vertices = their vertices, sorted by id;
  // hessian indices 0..n-1 over the non-fixed ones in that order.
  bool initializeOptimization(int level = 0) {
    HyperGraph::EdgeSet es;
    for (HyperGraph::EdgeSet::const_iterator it = edges_.begin(); it != edges_.end(); ++it)
      if (static_cast<OptimizableGraph::Edge*>(*it)->level() == level) es.insert(*it);
    return initializeOptimization(es);
  }
  bool initializeOptimization(HyperGraph::EdgeSet& eset) {
    active_edges_.clear();
    active_vertices_.clear();
    HyperGraph::VertexSet vs;
    for (HyperGraph::EdgeSet::const_iterator it = eset.begin(); it != eset.end(); ++it) {
      OptimizableGraph::Edge* e = static_cast<OptimizableGraph::Edge*>(*it);
      if (!e->vertex(0) || !e->vertex(1)) continue;
      active_edges_.push_back(e);
      vs.insert(e->vertex(0));
      vs.insert(e->vertex(1));
    }
    // A canonical order -- by end points, then measurement, then insertion -- instead of g2o's order of
    // insertion: LoopClosureChecker hands accepted closures over in heap-address order
    // (closure_checker.h:38), so the order of insertion differs from process to process, and with it
    // would the summation order of the solver and the spanning tree of computeInitialGuess (which
    // breaks hop-count ties by position in this list). With this order two processes holding the
    // same graph compute the same numbers.
    std::stable_sort(active_edges_.begin(), active_edges_.end(), canonical_before);
    for (VertexIDMap::iterator it = vertices_.begin(); it != vertices_.end(); ++it)
      static_cast<OptimizableGraph::Vertex*>(it->second)->setHessianIndex(-1);
    int h = 0;
    for (HyperGraph::VertexSet::iterator it = vs.begin(); it != vs.end(); ++it) {
      OptimizableGraph::Vertex* v = static_cast<OptimizableGraph::Vertex*>(*it);
      active_vertices_.push_back(v);
      v->setHessianIndex(v->fixed() ? -1 : h++);
    }
    structure_dirty_ = true;
    return !active_vertices_.empty();
  }

  static bool canonical_before(const OptimizableGraph::Edge* a, const OptimizableGraph::Edge* b) {
    const int a0 = a->vertex(0)->id(), a1 = a->vertex(1)->id(), b0 = b->vertex(0)->id(), b1 = b->vertex(1)->id();
    if (a0 != b0) return a0 < b0;
    if (a1 != b1) return a1 < b1;
    const EdgeSE2* ea = dynamic_cast<const EdgeSE2*>(a);
    const EdgeSE2* eb = dynamic_cast<const EdgeSE2*>(b);
    if (ea && eb) {
      const double ka[3] = {ea->measurement().translation().x(), ea->measurement().translation().y(),
                            ea->measurement().rotation().angle()};
      const double kb[3] = {eb->measurement().translation().x(), eb->measurement().translation().y(),
                            eb->measurement().rotation().angle()};
      for (int k = 0; k < 3; ++k)
        if (ka[k] != kb[k]) return ka[k] < kb[k];
    }
    return a->internalId() < b->internalId();
  }

  // C7: returns the iterations done (0 on solver failure)
  int optimize(int iterations) {
    if (!upload()) return 0;
    int done = 0;
    std::vector<double> poses(3 * active_vertices_.size());
    int rc;
    {
      trace::Clock clk(2);
      rc = pgo_iterate(solver_, iterations, poses.data(), nullptr, &done);
    }
    if (rc != PGO_OK) std::cerr << "optimize: " << pgo_last_error() << std::endl;
    if (done > 0) {
      trace_poses(trace::kOptimize, poses);
      download(poses);
    }
    if (verbose_) std::cerr << "optimize: " << done << " iterations" << std::endl;
    return rc == PGO_OK ? done : 0;
  }

  void computeInitialGuess() {  // C10
    if (!upload()) return;
    trace::Clock clk(4);
    if (pgo_initial_guess(solver_) != PGO_OK) {
      std::cerr << "computeInitialGuess: " << pgo_last_error() << std::endl;
      return;
    }
    std::vector<double> poses(3 * active_vertices_.size());
    if (pgo_get_poses(solver_, poses.data()) == PGO_OK) {
      trace_poses(trace::kInitialGuess, poses);
      download(poses);
    }
  }

  // C9: blockIndices are HESSIAN indices (as in graph_manipulator.cpp:134-142)
  bool computeMarginals(SparseBlockMatrix<Eigen::MatrixXd>& spinv,
                        const std::vector<std::pair<int, int> >& blockIndices) {
    if (!solver_ || structure_dirty_) return false;
    std::vector<int> by_hidx;
    for (size_t i = 0; i < active_vertices_.size(); ++i)
      if (active_vertices_[i]->hessianIndex() >= 0) by_hidx.push_back(static_cast<int>(i));
    std::vector<int32_t> r, c;
    for (size_t k = 0; k < blockIndices.size(); ++k) {
      const int hr = blockIndices[k].first, hc = blockIndices[k].second;
      if (hr < 0 || hc < 0 || hr >= static_cast<int>(by_hidx.size()) ||
          hc >= static_cast<int>(by_hidx.size()))
        return false;
      r.push_back(by_hidx[hr]);
      c.push_back(by_hidx[hc]);
    }
    std::vector<double> cov(9 * blockIndices.size());
    int rc_m;
    {
      trace::Clock clk(3);
      rc_m = pgo_marginals(solver_, static_cast<int>(r.size()), r.data(), c.data(), cov.data());
    }
    if (rc_m != PGO_OK) {
      std::cerr << "computeMarginals: " << pgo_last_error() << std::endl;
      return false;
    }
    if (trace::streams()[trace::current()].mode) {
      std::vector<long long> keys(blockIndices.size());
      for (size_t k = 0; k < blockIndices.size(); ++k)
        keys[k] = (static_cast<long long>(active_vertices_[r[k]]->id()) << 32) | static_cast<unsigned>(active_vertices_[c[k]]->id());
      trace::exchange(trace::kMarginals, keys, 9, cov, -1, true);
    }
    for (size_t k = 0; k < blockIndices.size(); ++k) {
      Eigen::MatrixXd* b = spinv.block(blockIndices[k].first, blockIndices[k].second, true);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) (*b)(i, j) = cov[9 * k + 3 * i + j];
    }
    return true;
  }

  // C14: g2o text format (VERTEX_SE2 / FIX / EDGE_SE2, level-0 edges)
  bool save(const char* filename, int level = 0) const {
    std::ofstream f(filename);
    if (!f) return false;
    f << std::setprecision(17);
    for (VertexIDMap::const_iterator it = vertices_.begin(); it != vertices_.end(); ++it) {
      const VertexSE2* v = dynamic_cast<const VertexSE2*>(it->second);
      if (!v) continue;
      f << "VERTEX_SE2 " << v->id() << " " << v->estimate().translation().x() << " "
        << v->estimate().translation().y() << " " << v->estimate().rotation().angle() << "\n";
    }
    for (VertexIDMap::const_iterator it = vertices_.begin(); it != vertices_.end(); ++it)
      if (static_cast<const OptimizableGraph::Vertex*>(it->second)->fixed())
        f << "FIX " << it->first << "\n";
    for (HyperGraph::EdgeSet::const_iterator it = edges_.begin(); it != edges_.end(); ++it) {
      const EdgeSE2* e = dynamic_cast<const EdgeSE2*>(*it);
      if (!e || e->level() != level) continue;
      const Eigen::Matrix3d& w = e->information();
      f << "EDGE_SE2 " << e->vertex(0)->id() << " " << e->vertex(1)->id() << " "
        << e->measurement().translation().x() << " " << e->measurement().translation().y() << " "
        << e->measurement().rotation().angle() << " " << w(0, 0) << " " << w(0, 1) << " "
        << w(0, 2) << " " << w(1, 1) << " " << w(1, 2) << " " << w(2, 2) << "\n";
    }
    return true;
  }
  bool load(const char* filename) {
    std::ifstream f(filename);
    if (!f) return false;
    std::string line, tag;
    while (std::getline(f, line)) {
      std::istringstream ss(line);
      if (!(ss >> tag)) continue;
      if (tag == "VERTEX_SE2") {
        int id;
        double x, y, th;
        ss >> id >> x >> y >> th;
        VertexSE2* v = new VertexSE2();
        v->setId(id);
        v->setEstimate(SE2(x, y, th));
        if (!addVertex(v)) delete v;
      } else if (tag == "FIX") {
        int id;
        while (ss >> id)
          if (vertex(id)) vertex(id)->setFixed(true);
      } else if (tag == "EDGE_SE2") {
        int a, b;
        double x, y, th, w[6];
        ss >> a >> b >> x >> y >> th;
        for (int k = 0; k < 6; ++k) ss >> w[k];
        if (!vertex(a) || !vertex(b)) continue;
        EdgeSE2* e = new EdgeSE2();
        e->vertices()[0] = vertex(a);
        e->vertices()[1] = vertex(b);
        e->setMeasurement(SE2(x, y, th));
        Eigen::Matrix3d m;
        m(0, 0) = w[0];
        m(0, 1) = m(1, 0) = w[1];
        m(0, 2) = m(2, 0) = w[2];
        m(1, 1) = w[3];
        m(1, 2) = m(2, 1) = w[4];
        m(2, 2) = w[5];
        e->setInformation(m);
        addEdge(e);
      }
    }
    return true;
  }

  pgo_solver* solver() const { return solver_; }
  // index of an active vertex in the arrays handed to the solver, or -1
  int activeIndex(const HyperGraph::Vertex* v) const {
    std::map<const HyperGraph::Vertex*, int>::const_iterator it = active_index_.find(v);
    return it == active_index_.end() ? -1 : it->second;
  }
  bool syncToDevice() { return upload(); }

 private:
  bool upload() {
    if (active_vertices_.empty()) {
      std::cerr << "SparseOptimizer: initializeOptimization has not been called" << std::endl;
      return false;
    }
    const int nv = static_cast<int>(active_vertices_.size());
    const int ne = static_cast<int>(active_edges_.size());
    // fixed flags may change between calls (GraphManipulator::fixGauge): they are structure
    std::vector<uint8_t> fixed(nv);
    for (int i = 0; i < nv; ++i) fixed[i] = active_vertices_[i]->fixed() ? 1 : 0;
    if (structure_dirty_ || fixed != last_fixed_ || !solver_) {
      active_index_.clear();
      for (int i = 0; i < nv; ++i) active_index_[active_vertices_[i]] = i;
      std::vector<int32_t> ei(ne), ej(ne);
      for (int e = 0; e < ne; ++e) {
        ei[e] = active_index_[active_edges_[e]->vertex(0)];
        ej[e] = active_index_[active_edges_[e]->vertex(1)];
      }
      // A keyframe alternates between two structures -- the graph with its own gauge
      // (graph_slam.cpp:392-393,564-565) and the same graph with the newest vertex as the gauge
      // (the covariance gate, graph_manipulator.cpp:116-145): one solver handle per structure, so
      // that the second optimize() of a keyframe finds its analysis and its captured iteration
      // graph still in place (pgo_set_graph recognises an unchanged structure).
      uint64_t key = 1469598103934665603ull;
      auto mix = [&key](uint64_t v) { key = (key ^ v) * 1099511628211ull; };
      mix(static_cast<uint64_t>(nv));
      mix(static_cast<uint64_t>(ne));
      for (int i = 0; i < nv; ++i) mix(fixed[i]);
      for (int e = 0; e < ne; ++e) mix((static_cast<uint64_t>(static_cast<uint32_t>(ei[e])) << 32) | static_cast<uint32_t>(ej[e]));
      int slot = -1;
      for (int k = 0; k < kSolverSlots; ++k)
        if (slots_[k] && slot_key_[k] == key) slot = k;
      if (slot < 0) {
        slot = 0;
        for (int k = 1; k < kSolverSlots; ++k)
          if (slot_used_[k] < slot_used_[slot]) slot = k;  // least recently used (or still empty)
      }
      if (!slots_[slot] && pgo_create(&slots_[slot], device_, nullptr) != PGO_OK) {
        std::cerr << "SparseOptimizer: " << pgo_last_error() << std::endl;
        slots_[slot] = nullptr;
        solver_ = nullptr;
        return false;
      }
      solver_ = slots_[slot];
      slot_key_[slot] = key;
      slot_used_[slot] = ++use_clock_;
      int rc_g;
      {
        trace::Clock clk(0);
        rc_g = pgo_set_graph(solver_, nv, ne, ei.data(), ej.data(), fixed.data());
      }
      if (rc_g != PGO_OK) {
        std::cerr << "SparseOptimizer: " << pgo_last_error() << std::endl;
        return false;
      }
      // hessian indices can change with the fixed flags
      int h = 0;
      for (int i = 0; i < nv; ++i) active_vertices_[i]->setHessianIndex(fixed[i] ? -1 : h++);
      last_fixed_ = fixed;
      structure_dirty_ = false;
    }
    std::vector<double> poses(3 * nv), meas(3 * ne), info(6 * ne);
    for (int i = 0; i < nv; ++i) {
      const SE2& x = static_cast<VertexSE2*>(active_vertices_[i])->estimate();
      poses[3 * i] = x.translation().x();
      poses[3 * i + 1] = x.translation().y();
      poses[3 * i + 2] = x.rotation().angle();
    }
    for (int e = 0; e < ne; ++e) {
      const EdgeSE2* ed = static_cast<const EdgeSE2*>(active_edges_[e]);
      meas[3 * e] = ed->measurement().translation().x();
      meas[3 * e + 1] = ed->measurement().translation().y();
      meas[3 * e + 2] = ed->measurement().rotation().angle();
      const Eigen::Matrix3d& w = ed->information();
      info[6 * e] = w(0, 0);
      info[6 * e + 1] = w(0, 1);
      info[6 * e + 2] = w(0, 2);
      info[6 * e + 3] = w(1, 1);
      info[6 * e + 4] = w(1, 2);
      info[6 * e + 5] = w(2, 2);
    }
    int rc_u;
    {
      trace::Clock clk(1);
      rc_u = pgo_upload(solver_, poses.data(), meas.data(), info.data());
    }
    if (rc_u != PGO_OK) {
      std::cerr << "SparseOptimizer: " << pgo_last_error() << std::endl;
      return false;
    }
    return true;
  }
  void trace_poses(int kind, std::vector<double>& poses) {
    if (!trace::streams()[trace::current()].mode) return;
    std::vector<long long> keys(active_vertices_.size());
    for (size_t i = 0; i < active_vertices_.size(); ++i) keys[i] = active_vertices_[i]->id();
    trace::exchange(kind, keys, 3, poses, 2, false);
  }
  void download(const std::vector<double>& poses) {
    for (size_t i = 0; i < active_vertices_.size(); ++i)
      if (!active_vertices_[i]->fixed())
        static_cast<VertexSE2*>(active_vertices_[i])
            ->setEstimate(SE2(poses[3 * i], poses[3 * i + 1], poses[3 * i + 2]));
  }

  VertexContainer active_vertices_;
  EdgeContainer active_edges_;
  std::map<const HyperGraph::Vertex*, int> active_index_;
  std::vector<uint8_t> last_fixed_;
  static const int kSolverSlots = 2;
  pgo_solver* slots_[kSolverSlots];
  uint64_t slot_key_[kSolverSlots];
  long long slot_used_[kSolverSlots];
  pgo_solver* solver_;  // the slot of the current structure
  OptimizationAlgorithm* algorithm_;
  bool verbose_;
  long long next_serial_;
  int device_;
  bool structure_dirty_;
  long long use_clock_;
};

// ---- C11: EdgeLabeler -----------------------------------------------------------------------------
class EdgeLabeler {
 public:
  explicit EdgeLabeler(SparseOptimizer* o) : opt_(o) {}
  // Star edges gauge -> v with the gauge fixed (condensed_graph_creator.cpp:49-63): sets each
  // edge's measurement from the current state and its information from the unscented transform.
  // Returns the number of edges labelled, -1 on failure. Assumes optimize() was just called.
  int labelEdges(std::set<OptimizableGraph::Edge*>& edges) {
    if (edges.empty()) return 0;
    if (!opt_->solver()) return -1;
    std::vector<EdgeSE2*> es;
    std::vector<int32_t> vs;
    int gauge = -1;
    for (std::set<OptimizableGraph::Edge*>::iterator it = edges.begin(); it != edges.end(); ++it) {
      EdgeSE2* e = dynamic_cast<EdgeSE2*>(*it);
      if (!e) return -1;
      OptimizableGraph::Vertex* a = static_cast<OptimizableGraph::Vertex*>(e->vertex(0));
      OptimizableGraph::Vertex* b = static_cast<OptimizableGraph::Vertex*>(e->vertex(1));
      if (!a->fixed() || b->fixed()) return -1;  // only the star pattern of the reference
      const int ga = opt_->activeIndex(a), vb = opt_->activeIndex(b);
      if (ga < 0 || vb < 0 || (gauge >= 0 && ga != gauge)) return -1;
      gauge = ga;
      es.push_back(e);
      vs.push_back(vb);
    }
    std::vector<double> meas(3 * es.size()), info(9 * es.size());
    int rc_l;
    {
      trace::Clock clk(5);
      rc_l = pgo_label_star_edges(opt_->solver(), gauge, static_cast<int>(es.size()), vs.data(), meas.data(),
                                  info.data());
    }
    if (rc_l != PGO_OK) {
      std::cerr << "labelEdges: " << pgo_last_error() << std::endl;
      return -1;
    }
    if (trace::streams()[trace::current()].mode) {
      std::vector<long long> keys(es.size());
      for (size_t k = 0; k < es.size(); ++k) keys[k] = es[k]->vertex(1)->id();
      trace::exchange(trace::kStarMeas, keys, 3, meas, 2, false);
      trace::exchange(trace::kStarInfo, keys, 9, info, -1, true);
    }
    for (size_t k = 0; k < es.size(); ++k) {
      es[k]->setMeasurement(SE2(meas[3 * k], meas[3 * k + 1], meas[3 * k + 2]));
      Eigen::Matrix3d m;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m(i, j) = info[9 * k + 3 * i + j];
      es[k]->setInformation(m);
    }
    return static_cast<int>(es.size());
  }

 private:
  SparseOptimizer* opt_;
};

}  // namespace g2o
#endif
