// C++ mirror of the reference's CharGrid (src/matcher/chargrid.h:106-231) over the C ABI of
// include/cgm_matcher.h: same type names, same member names, same argument meaning, so that
// src/matcher/scan_matcher.cpp-style callers compile against it. All arithmetic runs in
// libcgmrslam_b200.so (CUDA); this header only marshals. Failures are reported like the reference
// does: a line on std::cerr and an empty result.
//
// Differences a maintainer should know:
//   * a CharGrid is a HANDLE to device memory. Copying one (the reference copies CharGrid by value
//     in ScanMatcher::grid() and ScanMatcher::verifyMatching) shares the handle; clone() makes an
//     independent device copy of the cells.
//   * grid().cell(x, y) reads go through a host snapshot (CharGridMap::refresh()).
#ifndef CGM_CHARGRID_HPP
#define CGM_CHARGRID_HPP

#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <vector>

#include "../cgm_matcher.h"
#include "eigen_lite.hpp"

typedef std::vector<Eigen::Vector2i, Eigen::aligned_allocator<Eigen::Vector2i> > Vector2iVector;
typedef std::vector<Eigen::Vector2d, Eigen::aligned_allocator<Eigen::Vector2d> > Vector2dVector;

struct MatcherResult {  // chargrid.h:50-60
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
  Eigen::Vector3f lowerLeft;
  Eigen::Vector3f upperRight;
};
typedef std::vector<Region, Eigen::aligned_allocator<Region> > RegionVector;

struct MatchingParameters {  // chargrid.h:94-99
  Eigen::Vector3d searchStep;
  Eigen::Vector3d resultsDiscretization;
  double maxScore;
};
typedef std::vector<MatchingParameters, Eigen::aligned_allocator<MatchingParameters> >
    MatchingParametersVector;

namespace cgm {

struct Handle {
  cgm_matcher* m = nullptr;
  ~Handle() {
    if (m) cgm_matcher_destroy(m);
  }
};

inline bool ok(int rc, const char* what) {
  if (rc == CGM_OK) return true;
  std::cerr << "cgm: " << what << " failed: " << cgm_last_error() << std::endl;
  return false;
}

}  // namespace cgm

// The subset of _GridMap<unsigned char> (gridmap.h) that callers of CharGrid use.
class CharGridMap {
 public:
  CharGridMap() {}
  explicit CharGridMap(std::shared_ptr<cgm::Handle> h, float res) : h_(h), res_(res) {}
  Eigen::Vector2i size() const {
    int r = 0, c = 0;
    if (h_) cgm_matcher_grid_size(h_->m, &r, &c);
    return Eigen::Vector2i(r, c);
  }
  float resolution() const { return res_; }
  float inverseResolution() const { return static_cast<float>(1. / res_); }
  Eigen::Vector2i world2grid(const Eigen::Vector2f& wp) const {
    int x = 0, y = 0;
    cgm_matcher_world2grid(h_->m, wp.x(), wp.y(), &x, &y);
    return Eigen::Vector2i(x, y);
  }
  Eigen::Vector2f grid2world(const Eigen::Vector2i& gp) const {
    float x = 0, y = 0;
    cgm_matcher_grid2world(h_->m, gp.x(), gp.y(), &x, &y);
    return Eigen::Vector2f(x, y);
  }
  bool isInside(const Eigen::Vector2i& mp) const {
    const Eigen::Vector2i s = size();
    return mp.x() >= 0 && mp.y() >= 0 && mp.x() < s.x() && mp.y() < s.y();
  }
  // Host snapshot of the cells (row-major [x][y]) for read access.
  void refresh() const {
    const Eigen::Vector2i s = size();
    cells_.resize(static_cast<size_t>(s.x()) * s.y());
    cols_ = s.y();
    if (!cells_.empty()) cgm::ok(cgm_matcher_grid_download(h_->m, 0, cells_.data()), "grid download");
  }
  unsigned char cell(int x, int y) const {
    if (cells_.empty()) refresh();
    return cells_[static_cast<size_t>(x) * cols_ + y];
  }
  unsigned char cell(const Eigen::Vector2i& p) const { return cell(p.x(), p.y()); }

 private:
  std::shared_ptr<cgm::Handle> h_;
  float res_ = 0.f;
  mutable std::vector<unsigned char> cells_;
  mutable int cols_ = 0;
};

struct CharGrid {
  // chargrid.cpp:124-128. kernelRange / kernelResolution configure the stamp that
  // addAndConvolvePoints applies (the reference passes the MatrixXChar built by
  // ScanMatcher::initializeKernel; here the library builds the same stamp from the same two
  // numbers, scan_matcher.cpp:38-61).
  CharGrid() {}
  CharGrid(Eigen::Vector2f lowerLeft_, Eigen::Vector2f upperRight_, float res_, int kscale_ = 128,
           double kernelResolution = 0.0, double kernelRange = 0.0, int device = 0)
      : h_(new cgm::Handle()), ll_(lowerLeft_), ur_(upperRight_), res_(res_), kscale_(kscale_),
        kres_(kernelResolution > 0.0 ? kernelResolution : res_), krange_(kernelRange),
        device_(device) {
    cgm::ok(cgm_matcher_create(&h_->m, device, nullptr, 1, lowerLeft_.x(), lowerLeft_.y(),
                               upperRight_.x(), upperRight_.y(), kres_, kernelRange, kscale_),
            "CharGrid");
    // the grid itself is built from the float resolution (scan_matcher.cpp:64); when the two
    // resolutions differ the caller must use matching doubles/floats as the reference does
    map_ = CharGridMap(h_, res_);
  }
  bool valid() const { return h_ && h_->m; }
  cgm_matcher* handle() const { return h_ ? h_->m : nullptr; }

  // independent device copy (what `CharGrid auxGrid = _grid;` means in the reference)
  CharGrid clone() const {
    CharGrid c(ll_, ur_, res_, kscale_, kres_, krange_, device_);
    if (valid() && c.valid()) {
      map_.refresh();
      const Eigen::Vector2i s = map_.size();
      std::vector<unsigned char> cells(static_cast<size_t>(s.x()) * s.y());
      cgm::ok(cgm_matcher_grid_download(h_->m, 0, cells.data()), "clone download");
      cgm::ok(cgm_matcher_grid_upload(c.h_->m, 0, cells.data()), "clone upload");
    }
    return c;
  }

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

  static void subsample(Vector2dVector& dest, const Vector2dVector& src, double res) {
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

  const CharGridMap& grid() const { return map_; }
  CharGridMap& grid() { return map_; }

  // ScanMatcher::resetGrid (scan_matcher.cpp:68-76): every cell <- int(kernelRange * kscale)
  void reset() { cgm::ok(cgm_matcher_reset(h_->m, 0), "reset"); }

  template <typename T>
  void addAndConvolvePoints(typename T::const_iterator begin_, typename T::const_iterator end_) {
    std::vector<double> xy;
    for (typename T::const_iterator it = begin_; it != end_; ++it) {
      xy.push_back(it->x());
      xy.push_back(it->y());
    }
    cgm::ok(cgm_matcher_raster(h_->m, 0, xy.data(), static_cast<int>(xy.size() / 2)),
            "addAndConvolvePoints");
  }
  // signature-compatible overload: the kernel argument is ignored (the handle owns the stamp)
  template <typename T, typename K>
  void addAndConvolvePoints(typename T::const_iterator begin_, typename T::const_iterator end_,
                            const K&) {
    addAndConvolvePoints<T>(begin_, end_);
  }

  double greedySearch(Eigen::Vector3d& result, const Vector2dVector& points,
                      Eigen::Vector3f lowerLeftF, Eigen::Vector3f upperRightF, double thetaRes,
                      double maxScore) {  // chargrid.cpp:163-180
    std::vector<MatcherResult> mresvec;
    const double dx = res_ * 4, dy = res_ * 4, dth = thetaRes * 4;
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
    params.searchStep = Eigen::Vector3d(res_, res_, thetaRes);
    params.maxScore = maxScore;
    params.resultsDiscretization = Eigen::Vector3d(dx, dy, dth);
    greedySearch(mresvec, points, regions, params);
  }

  void greedySearch(std::vector<MatcherResult>& mresvec, const Vector2dVector& points,
                    const RegionVector& regions, const MatchingParameters& params) {  // :208-308
    mresvec.clear();
    std::vector<double> xy;
    std::vector<float> reg;
    pack(points, regions, &xy, &reg);
    int cap = 256, n = 0;
    std::vector<cgm_result> out(cap);
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (!cgm::ok(cgm_matcher_search(h_->m, 0, xy.data(), static_cast<int>(points.size()),
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
                          const RegionVector& regions, double thetaRes, double maxScore, double dx,
                          double dy, double dth, int nLevels) {  // chargrid.cpp:376-400
    mresvec.clear();
    std::vector<double> xy;
    std::vector<float> reg;
    pack(points, regions, &xy, &reg);
    int cap = 1024, n = 0;
    std::vector<cgm_result> out(cap);
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (!cgm::ok(cgm_matcher_hierarchical_search(
                       h_->m, 0, xy.data(), static_cast<int>(points.size()), reg.data(),
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
    cgm::ok(cgm_matcher_count_points(h_->m, 0, lowerLeftF.x(), lowerLeftF.y(), upperRightF.x(),
                                     upperRightF.y(), score),
            "countPoints");
  }

  void searchNonMatchedPoints(const Vector2dVector& points, Vector2dVector& nonmatchedpoints,
                              double maxScore) {
    std::vector<double> xy(2 * points.size()), out(2 * points.size());
    for (size_t i = 0; i < points.size(); ++i) {
      xy[2 * i] = points[i].x();
      xy[2 * i + 1] = points[i].y();
    }
    int n = 0;
    nonmatchedpoints.clear();
    if (!cgm::ok(cgm_matcher_search_non_matched(h_->m, 0, xy.data(), static_cast<int>(points.size()),
                                                maxScore, out.data(), &n),
                 "searchNonMatchedPoints"))
      return;
    for (int i = 0; i < n; ++i) nonmatchedpoints.push_back(Eigen::Vector2d(out[2 * i], out[2 * i + 1]));
  }

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

  std::shared_ptr<cgm::Handle> h_;
  Eigen::Vector2f ll_, ur_;
  float res_ = 0.f;
  int kscale_ = 128;
  double kres_ = 0.0, krange_ = 0.0;
  int device_ = 0;
  CharGridMap map_;
};

#endif
