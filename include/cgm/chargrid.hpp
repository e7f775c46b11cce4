// Drop-in replacement for the reference's src/matcher/chargrid.h + gridmap.h (+ chargrid.cpp): the
// same type names, member names, signatures and value semantics, so that the reference's own
// src/matcher/scan_matcher.cpp (and everything above it) compiles and links UNCHANGED against it:
//
//     g++ -I include/ref_names -include include/cgm/chargrid.hpp ... src/matcher/scan_matcher.cpp
//
// This header defines the reference headers' include guards, so their `#include "chargrid.h"`
// (resolved next to the including file, which -I cannot override) contributes nothing.
// All arithmetic runs in libcgmrslam_b200.so (CUDA) through include/cgm_matcher.h; this header only
// marshals. Failures are reported like the reference does: a line on std::cerr and an empty result.
//
// What a maintainer should know:
//   * A CharGrid owns one device grid. Copy construction / assignment are deep (device-to-device),
//     as in the reference (`CharGrid auxGrid = _grid;`, scan_matcher.cpp:437; `_grid = tmp;`, :65).
//   * grid().cell(x, y) is a reference into a HOST mirror of the cells. The mirror is downloaded on
//     first touch after the device wrote the grid and written back before the device reads it. A
//     mirror holding one value everywhere -- the state ScanMatcher::resetGrid leaves behind
//     (scan_matcher.cpp:68-76) -- is not uploaded: the device fills the grid itself, in the same
//     launch that stamps the points of the following addAndConvolvePoints.
#ifndef CGM_CHARGRID_HPP
#define CGM_CHARGRID_HPP
#ifndef _CHARGRID_MAP_H_
#define _CHARGRID_MAP_H_  // src/matcher/chargrid.h
#endif
#ifndef GRIDMAP_HH
#define GRIDMAP_HH  // src/matcher/gridmap.h
#endif

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <vector>

#include "../cgm_matcher.h"
#include "device.hpp"
#include "eigen_lite.hpp"

typedef std::vector<Eigen::Vector2i, Eigen::aligned_allocator<Eigen::Vector2i> > Vector2iVector;
typedef std::vector<Eigen::Vector2d, Eigen::aligned_allocator<Eigen::Vector2d> > Vector2dVector;
typedef Eigen::Matrix<unsigned char, Eigen::Dynamic, Eigen::Dynamic> MatrixXChar;

struct MatcherResult {  // chargrid.h:50-60
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
  MatcherResult(const Eigen::Vector3d& transformation_, const double& score_,
                const Eigen::Matrix3d& informationMatrix_ = Eigen::Matrix3d::Identity())
      : transformation(transformation_), score(score_), informationMatrix(informationMatrix_) {}
  Eigen::Vector3d transformation;
  double score;
  Eigen::Matrix3d informationMatrix;
};

struct MatcherResultScoreComparator {  // chargrid.h:62-66
  bool operator()(const MatcherResult& a, const MatcherResult& b) const { return a.score < b.score; }
};

struct DiscreteTriplet {  // chargrid.h:68-85
  DiscreteTriplet(const Eigen::Vector3d& tr, double dx, double dy, double dth) {
    ix = static_cast<int>(tr.x() / dx);
    iy = static_cast<int>(tr.y() / dy);
    ith = static_cast<int>(tr.z() / dth);
  }
  bool operator<(const DiscreteTriplet& o) const {
    if (ix < o.ix) return true;
    if (ix == o.ix && iy < o.iy) return true;
    if (ix == o.ix && iy == o.iy && ith < o.ith) return true;
    return false;
  }
  double ix, iy, ith;
};

struct Region {  // chargrid.h:87-92
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
  Eigen::Vector3f lowerLeft;
  Eigen::Vector3f upperRight;
};
typedef std::vector<Region, Eigen::aligned_allocator<Region> > RegionVector;

struct MatchingParameters {  // chargrid.h:94-99
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
  Eigen::Vector3d searchStep;
  Eigen::Vector3d resultsDiscretization;
  double maxScore;
};
typedef std::vector<MatchingParameters, Eigen::aligned_allocator<MatchingParameters> >
    MatchingParametersVector;

namespace cgm {

inline bool ok(int rc, const char* what) {
  if (rc == CGM_OK) return true;
  std::cerr << "cgm: " << what << " failed: " << cgm_last_error() << std::endl;
  return false;
}

}  // namespace cgm

// The part of _GridMap<unsigned char> (gridmap.h) that is reachable through CharGrid: geometry,
// cell access, copy / assignment. Owns the device grid and its host mirror.
class CharGridMap {
 public:
  typedef unsigned char CellType;

  CharGridMap() {}
  CharGridMap(const Eigen::Vector2f& lowerLeft, const Eigen::Vector2f& upperRight, float res, int kscale)
      : ll_(lowerLeft), ur_(upperRight), res_(res), kscale_(kscale) {
    create();
  }
  CharGridMap(const CharGridMap& o) : ll_(o.ll_), ur_(o.ur_), res_(o.res_), kscale_(o.kscale_) {
    if (o.m_) {
      create();
      copyCells(o);
    }
  }
  CharGridMap& operator=(const CharGridMap& o) {
    if (this == &o) return *this;
    const bool same = m_ && o.m_ && rows_ == o.rows_ && cols_ == o.cols_ && ll_.x() == o.ll_.x() &&
                      ll_.y() == o.ll_.y() && res_ == o.res_ && kscale_ == o.kscale_;
    if (!same) {
      destroy();
      ll_ = o.ll_;
      ur_ = o.ur_;
      res_ = o.res_;
      kscale_ = o.kscale_;
      if (o.m_) create();
    }
    if (o.m_) copyCells(o);
    return *this;
  }
  ~CharGridMap() { destroy(); }

  // ---- gridmap.h geometry -------------------------------------------------------------------------
  Eigen::Vector2i size() const { return Eigen::Vector2i(rows_, cols_); }
  float resolution() const { return res_; }
  float inverseResolution() const { return static_cast<float>(1. / res_); }
  const Eigen::Vector2f& lowerLeft() const { return ll_; }
  const Eigen::Vector2f& upperRight() const { return ur_; }
  Eigen::Vector2f center() const { return Eigen::Vector2f(.5 * (ur_.x() + ll_.x()), .5 * (ur_.y() + ll_.y())); }
  Eigen::Vector2i world2grid(const Eigen::Vector2f& wp) const { return world2grid(wp.x(), wp.y()); }
  Eigen::Vector2i world2grid(const float x, const float y) const {
    int ix = 0, iy = 0;
    if (m_) cgm_matcher_world2grid(m_, x, y, &ix, &iy);
    return Eigen::Vector2i(ix, iy);
  }
  Eigen::Vector2f grid2world(const Eigen::Vector2i& gp) const {
    float x = 0, y = 0;
    if (m_) cgm_matcher_grid2world(m_, gp.x(), gp.y(), &x, &y);
    return Eigen::Vector2f(x, y);
  }
  bool isInside(const Eigen::Vector2i& mp) const {
    return mp.x() >= 0 && mp.y() >= 0 && mp.x() < rows_ && mp.y() < cols_;
  }

  // ---- cell access: references into the host mirror ----------------------------------------------
  // ScanMatcher::resetGrid writes all 1.44 M cells through this accessor (scan_matcher.cpp:71-74), and a
  // store through unsigned char& may alias every member, so the compiler reloads them per call: the
  // flags are one byte that is only READ on the common path (a per-call store to a flag next to the
  // load of the other costs a store-forwarding stall: 2.9 -> 1.3 ms for the reset loop).
  unsigned char& cell(const int& x, const int& y) {
    if (flags_ != 3) {
      touch();
      set_dirty(true);
    }
    return cells_[static_cast<size_t>(x) * cols_ + y];
  }
  const unsigned char& cell(const int& x, const int& y) const {
    if (!(flags_ & 1)) touch();
    return cells_[static_cast<size_t>(x) * cols_ + y];
  }
  unsigned char& cell(const Eigen::Vector2i& p) { return cell(p.x(), p.y()); }
  const unsigned char& cell(const Eigen::Vector2i& p) const { return cell(p.x(), p.y()); }

  void saveAsPPM(std::ostream& os, bool equalize) const {  // gridmap.h: debugging aid
    touch();
    os << "P6\n" << rows_ << " " << cols_ << "\n255\n";
    int hi = 1;
    if (equalize)
      for (size_t i = 0; i < cells_.size(); ++i) hi = cells_[i] > hi ? cells_[i] : hi;
    for (int y = cols_ - 1; y >= 0; --y)
      for (int x = 0; x < rows_; ++x) {
        const unsigned char c = equalize ? static_cast<unsigned char>(255 * cells_[static_cast<size_t>(x) * cols_ + y] / hi)
                                         : cells_[static_cast<size_t>(x) * cols_ + y];
        os.put(c).put(c).put(c);
      }
  }

  // ---- device side (used by CharGrid) ----------------------------------------------------------------
  cgm_matcher* handle() const { return m_; }
  bool valid() const { return m_ != nullptr; }
  // make the device grid current before a kernel reads it
  void flush() {
    if (!m_ || !host_dirty()) return;
    int value = 0;
    if (uniform(&value)) cgm::ok(cgm_matcher_fill(m_, 0, value), "grid fill");
    else cgm::ok(cgm_matcher_grid_upload(m_, 0, cells_.data()), "grid upload");
    set_dirty(false);
  }
  // CharGrid::applyKernel over a list of world points (chargrid.h:205-216, chargrid.cpp:132-161)
  void stamp(const MatrixXChar& kernel, const std::vector<double>& xy) {
    if (!m_) return;
    useKernel(kernel);
    const int n = static_cast<int>(xy.size() / 2);
    int value = 0;
    if (host_dirty() && uniform(&value)) {  // resetGrid + addAndConvolvePoints: one launch
      cgm::ok(cgm_matcher_fill_raster(m_, 0, value, xy.data(), n), "addAndConvolvePoints");
      set_dirty(false);
    } else {
      flush();
      cgm::ok(cgm_matcher_raster(m_, 0, xy.data(), n), "addAndConvolvePoints");
    }
    set_valid(false);
  }

 private:
  void create() {
    if (!cgm::ok(cgm_matcher_create(&m_, cgm::default_device(), nullptr, 1, ll_.x(), ll_.y(), ur_.x(), ur_.y(),
                                    res_, 0.0, kscale_),
                 "CharGrid")) {
      m_ = nullptr;
      return;
    }
    cgm_matcher_grid_size(m_, &rows_, &cols_);
    // gridmap.h:196-214 value-initialises the cells
    cells_.assign(static_cast<size_t>(rows_) * cols_, 0);
    set_valid(true);
    set_dirty(true);
    stamp_dim_ = -1;
  }
  void destroy() {
    if (m_) cgm_matcher_destroy(m_);
    m_ = nullptr;
    rows_ = cols_ = 0;
    cells_.clear();
  }
  void copyCells(const CharGridMap& o) {
    if (!m_) return;
    if (o.host_dirty()) {  // the host mirror of `o` is the newer copy
      cells_ = o.cells_;
      set_valid(true);
      set_dirty(true);
    } else {
      cgm::ok(cgm_matcher_copy_grid(m_, 0, o.m_, 0), "CharGrid copy");
      set_valid(o.host_valid());
      if (host_valid()) cells_ = o.cells_;
      set_dirty(false);
    }
  }
  void touch() const {
    if (host_valid() || !m_) return;
    cells_.resize(static_cast<size_t>(rows_) * cols_);
    cgm::ok(cgm_matcher_grid_download(m_, 0, cells_.data()), "grid download");
    set_valid(true);
  }
  bool uniform(int* value) const {
    if (cells_.empty()) return false;
    const unsigned char v = cells_[0];
    *value = v;
    return cells_.size() == 1 || std::memcmp(cells_.data(), cells_.data() + 1, cells_.size() - 1) == 0;
  }
  void useKernel(const MatrixXChar& kernel) {
    const int dim = static_cast<int>(kernel.rows());
    const size_t bytes = static_cast<size_t>(kernel.rows()) * kernel.cols();
    if (dim == stamp_dim_ && bytes == stamp_.size() && std::memcmp(stamp_.data(), kernel.data(), bytes) == 0)
      return;
    if (kernel.rows() != kernel.cols() ||
        !cgm::ok(cgm_matcher_set_stamp(m_, kernel.data(), dim), "kernel")) {
      stamp_dim_ = -1;
      return;
    }
    stamp_.assign(kernel.data(), kernel.data() + bytes);
    stamp_dim_ = dim;
  }

  cgm_matcher* m_ = nullptr;
  Eigen::Vector2f ll_, ur_;
  float res_ = 0.f;
  int kscale_ = 128;
  int rows_ = 0, cols_ = 0;
  mutable std::vector<unsigned char> cells_;
  // bit 0: the mirror equals the device grid (or is newer: bit 1); bit 1: the mirror holds writes the
  // device has not seen. One byte, so that cell() tests both with one load.
  mutable unsigned char flags_ = 0;
  bool host_valid() const { return (flags_ & 1) != 0; }
  bool host_dirty() const { return (flags_ & 2) != 0; }
  void set_valid(bool v) const { flags_ = static_cast<unsigned char>(v ? (flags_ | 1) : (flags_ & ~1)); }
  void set_dirty(bool v) { flags_ = static_cast<unsigned char>(v ? (flags_ | 2) : (flags_ & ~2)); }
  std::vector<unsigned char> stamp_;
  int stamp_dim_ = -1;
};

struct CharGrid {  // chargrid.h:106-231
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW;

  CharGrid(Eigen::Vector2f lowerLeft_, Eigen::Vector2f upperRight_, float res_, int kscale_ = 128)
      : _grid(lowerLeft_, upperRight_, res_, kscale_), _kscale(kscale_) {}
  ~CharGrid() {}

  static void addToPrunedMap(std::map<DiscreteTriplet, MatcherResult>& myMap, MatcherResult& mr,
                             double dx, double dy, double dth) {  // chargrid.cpp:36-46
    DiscreteTriplet t(mr.transformation, dx, dy, dth);
    std::map<DiscreteTriplet, MatcherResult>::iterator it = myMap.find(t);
    if (it != myMap.end()) {
      if (it->second.score > mr.score) it->second = mr;
    } else {
      myMap.insert(std::make_pair(t, mr));
    }
  }

  static void subsample(Vector2dVector& dest, const Vector2dVector& src, double res) {  // :98-122
    std::vector<double> in(2 * src.size()), out(2 * src.size());
    for (size_t i = 0; i < src.size(); ++i) {
      in[2 * i] = src[i].x();
      in[2 * i + 1] = src[i].y();
    }
    int n = 0;
    cgm::ok(cgm_subsample(in.data(), static_cast<int>(src.size()), res, out.data(), &n), "subsample");
    dest.resize(n);
    for (int i = 0; i < n; ++i) dest[i] = Eigen::Vector2d(out[2 * i], out[2 * i + 1]);
  }

  inline const CharGridMap& grid() const { return _grid; }
  inline CharGridMap& grid() { return _grid; }

  double greedySearch(Eigen::Vector3d& result, const Vector2dVector& points,
                      Eigen::Vector3f lowerLeftF, Eigen::Vector3f upperRightF, double thetaRes,
                      double maxScore) {  // chargrid.cpp:163-180
    std::vector<MatcherResult> mresvec;
    const double dx = _grid.resolution() * 4, dy = _grid.resolution() * 4, dth = thetaRes * 4;
    greedySearch(mresvec, points, lowerLeftF, upperRightF, thetaRes, maxScore, dx, dy, dth);
    if (mresvec.size()) {
      result = mresvec[0].transformation;
      return mresvec[0].score;
    }
    return std::numeric_limits<double>::max();
  }

  void greedySearch(std::vector<MatcherResult>& mresvec, const Vector2dVector& points,
                    Eigen::Vector3f lowerLeftF, Eigen::Vector3f upperRightF, double thetaRes,
                    double maxScore, double dx, double dy, double dth) {  // chargrid.cpp:183-194
    RegionVector regions(1);
    regions[0].lowerLeft = lowerLeftF;
    regions[0].upperRight = upperRightF;
    greedySearch(mresvec, points, regions, thetaRes, maxScore, dx, dy, dth);
  }

  void greedySearch(std::vector<MatcherResult>& mresvec, const Vector2dVector& points,
                    const RegionVector& regions, double thetaRes, double maxScore, double dx,
                    double dy, double dth) {  // chargrid.cpp:196-206
    MatchingParameters params;
    params.searchStep = Eigen::Vector3d(_grid.resolution(), _grid.resolution(), thetaRes);
    params.maxScore = maxScore;
    params.resultsDiscretization = Eigen::Vector3d(dx, dy, dth);
    greedySearch(mresvec, points, regions, params);
  }

  void greedySearch(std::vector<MatcherResult>& mresvec, const Vector2dVector& points,
                    const RegionVector& regions, const MatchingParameters& params) {  // :208-308
    mresvec.clear();
    if (!_grid.valid()) return;
    _grid.flush();
    std::vector<double> xy;
    std::vector<float> reg;
    pack(points, regions, &xy, &reg);
    int cap = 256, n = 0;
    std::vector<cgm_result> out(cap);
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (!cgm::ok(cgm_matcher_search(_grid.handle(), 0, xy.data(), static_cast<int>(points.size()),
                                      reg.data(), static_cast<int>(regions.size()),
                                      params.searchStep.x(), params.searchStep.y(),
                                      params.searchStep.z(), params.maxScore,
                                      params.resultsDiscretization.x(),
                                      params.resultsDiscretization.y(),
                                      params.resultsDiscretization.z(), out.data(), cap, &n),
                   "greedySearch"))
        return;
      if (n <= cap) break;
      cap = n;
      out.resize(cap);
    }
    unpack(out, n, &mresvec);
  }

  void hierarchicalSearch(std::vector<MatcherResult>& mresvec, const Vector2dVector& points,
                          const RegionVector& regions,
                          const MatchingParametersVector& paramsVec) {  // chargrid.cpp:310-344
    mresvec.clear();
    if (!_grid.valid() || paramsVec.empty()) return;
    _grid.flush();
    std::vector<double> xy, levels(7 * paramsVec.size());
    std::vector<float> reg;
    pack(points, regions, &xy, &reg);
    for (size_t l = 0; l < paramsVec.size(); ++l) {
      for (int c = 0; c < 3; ++c) {
        levels[7 * l + c] = paramsVec[l].searchStep[c];
        levels[7 * l + 4 + c] = paramsVec[l].resultsDiscretization[c];
      }
      levels[7 * l + 3] = paramsVec[l].maxScore;
    }
    int cap = 1024, n = 0;
    std::vector<cgm_result> out(cap);
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (!cgm::ok(cgm_matcher_hierarchical_search_levels(
                       _grid.handle(), 0, xy.data(), static_cast<int>(points.size()), reg.data(),
                       static_cast<int>(regions.size()), levels.data(), static_cast<int>(paramsVec.size()),
                       out.data(), cap, &n),
                   "hierarchicalSearch"))
        return;
      if (n <= cap) break;
      cap = n;
      out.resize(cap);
    }
    unpack(out, n, &mresvec);
  }

  void hierarchicalSearch(std::vector<MatcherResult>& mresvec, const Vector2dVector& points,
                          const RegionVector& regions, double thetaRes, double maxScore, double dx,
                          double dy, double dth, int nLevels) {  // chargrid.cpp:376-400
    mresvec.clear();
    if (!_grid.valid()) return;
    _grid.flush();
    std::vector<double> xy;
    std::vector<float> reg;
    pack(points, regions, &xy, &reg);
    int cap = 1024, n = 0;
    std::vector<cgm_result> out(cap);
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (!cgm::ok(cgm_matcher_hierarchical_search(
                       _grid.handle(), 0, xy.data(), static_cast<int>(points.size()), reg.data(),
                       static_cast<int>(regions.size()), thetaRes, maxScore, dx, dy, dth, nLevels,
                       out.data(), cap, &n),
                   "hierarchicalSearch"))
        return;
      if (n <= cap) break;
      cap = n;
      out.resize(cap);
    }
    unpack(out, n, &mresvec);
  }

  void hierarchicalSearch(std::vector<MatcherResult>& mresvec, const Vector2dVector& points,
                          Eigen::Vector3f lowerLeftF, Eigen::Vector3f upperRightF, double thetaRes,
                          double maxScore, double dx, double dy, double dth, int nLevels) {
    RegionVector regions(1);  // chargrid.cpp:402-414
    regions[0].lowerLeft = lowerLeftF;
    regions[0].upperRight = upperRightF;
    hierarchicalSearch(mresvec, points, regions, thetaRes, maxScore, dx, dy, dth, nLevels);
  }

  void hierarchicalSearch(std::vector<MatcherResult>& mresvec, const Vector2dVector& points,
                          Eigen::Vector3f lowerLeftF, Eigen::Vector3f upperRightF, double thetaRes,
                          double maxScore, double dx, double dy, double dth) {
    // chargrid.cpp:346-374 hard-codes the same ladder as nLevels = 4
    hierarchicalSearch(mresvec, points, lowerLeftF, upperRightF, thetaRes, maxScore, dx, dy, dth, 4);
  }

  void countPoints(Eigen::Vector2f lowerLeftF, Eigen::Vector2f upperRightF, double* score) {
    if (!_grid.valid()) return;
    _grid.flush();
    cgm::ok(cgm_matcher_count_points(_grid.handle(), 0, lowerLeftF.x(), lowerLeftF.y(), upperRightF.x(),
                                     upperRightF.y(), score),
            "countPoints");
  }

  void searchNonMatchedPoints(const Vector2dVector& points, Vector2dVector& nonmatchedpoints,
                              double maxScore) {
    nonmatchedpoints.clear();
    if (!_grid.valid()) return;
    _grid.flush();
    std::vector<double> xy(2 * points.size()), out(2 * points.size());
    for (size_t i = 0; i < points.size(); ++i) {
      xy[2 * i] = points[i].x();
      xy[2 * i + 1] = points[i].y();
    }
    int n = 0;
    if (!cgm::ok(cgm_matcher_search_non_matched(_grid.handle(), 0, xy.data(), static_cast<int>(points.size()),
                                                maxScore, out.data(), &n),
                 "searchNonMatchedPoints"))
      return;
    for (int i = 0; i < n; ++i) nonmatchedpoints.push_back(Eigen::Vector2d(out[2 * i], out[2 * i + 1]));
  }

  // chargrid.cpp:132-161: min-stamp `kernel` centred on cell (r, c) of `out`. grid2world of a cell
  // maps back to that cell under world2grid (|rounding| << 0.5 cell), so the point path serves it.
  void applyKernel(CharGridMap& out, const MatrixXChar& kernel, int r, int c) {
    const Eigen::Vector2f p = out.grid2world(Eigen::Vector2i(r, c));
    std::vector<double> xy(2);
    xy[0] = p.x();
    xy[1] = p.y();
    out.stamp(kernel, xy);
  }

  template <typename T>
  void addPoints(typename T::const_iterator begin_, typename T::const_iterator end_, const float& val = 0.) {
    for (typename T::const_iterator it = begin_; it != end_; ++it) {  // chargrid.h:192-203
      Eigen::Vector2d myP = *it;
      Eigen::Vector2f p(myP.x(), myP.y());
      Eigen::Vector2i ip = _grid.world2grid(p);
      if (_grid.isInside(ip)) _grid.cell(ip) = val;  // (the reference writes out of bounds here)
    }
  }

  template <typename T>
  void addAndConvolvePoints(typename T::const_iterator begin_, typename T::const_iterator end_,
                            const MatrixXChar& kernel) {  // chargrid.h:205-216
    std::vector<double> xy;
    for (typename T::const_iterator it = begin_; it != end_; ++it) {
      xy.push_back(it->x());
      xy.push_back(it->y());
    }
    _grid.stamp(kernel, xy);
  }

  template <typename T>
  void integrateScan(T& m, typename T::const_iterator begin_, typename T::const_iterator end_,
                     const Eigen::Isometry2d& tsf) {  // chargrid.h:218-227
    for (typename T::const_iterator it = begin_; it != end_; ++it) {
      Eigen::Vector2d p = *it;
      p = tsf * p;
      m.push_back(p);
    }
  }

  CharGridMap _grid;
  int _kscale;

 private:
  static void pack(const Vector2dVector& points, const RegionVector& regions,
                   std::vector<double>* xy, std::vector<float>* reg) {
    xy->resize(2 * points.size());
    for (size_t i = 0; i < points.size(); ++i) {
      (*xy)[2 * i] = points[i].x();
      (*xy)[2 * i + 1] = points[i].y();
    }
    reg->resize(6 * regions.size());
    for (size_t i = 0; i < regions.size(); ++i)
      for (int c = 0; c < 3; ++c) {
        (*reg)[6 * i + c] = regions[i].lowerLeft[c];
        (*reg)[6 * i + 3 + c] = regions[i].upperRight[c];
      }
  }
  static void unpack(const std::vector<cgm_result>& out, int n, std::vector<MatcherResult>* res) {
    for (int i = 0; i < n; ++i)
      res->push_back(MatcherResult(Eigen::Vector3d(out[i].x, out[i].y, out[i].theta), out[i].score));
  }
};

#endif
