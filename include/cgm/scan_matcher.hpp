// C++ mirror of the reference's ScanMatcher (src/matcher/scan_matcher.h:41-85,
// src/matcher/scan_matcher.cpp): same method names, argument meaning and error behaviour
// (bool found / not found, diagnostics on stderr). The front half -- gathering the scans of a
// vertex set in the reference vertex's frame, window set-up, result -> SE2 -- is host code that
// follows scan_matcher.cpp line by line in *behaviour*; rasterisation and every search run on the
// GPU through CharGrid (include/cgm/chargrid.hpp -> include/cgm_matcher.h).
#ifndef CGM_SCAN_MATCHER_HPP
#define CGM_SCAN_MATCHER_HPP

#include <cstdio>
#include <ctime>
#include <iostream>
#include <map>
#include <vector>

#include "../g2o_compat/g2o_compat.hpp"
#include "chargrid.hpp"

namespace cgm {

// scan_matcher.cpp:89-110: every vertex's RobotLaser::cartesian() mapped into the reference
// vertex's frame by (ref^-1 * v) * laserPose; the reference vertex itself only by laserPose.
inline void transformPointsFromVSet(g2o::OptimizableGraph::VertexSet& vset,
                                    g2o::OptimizableGraph::Vertex* reference_,
                                    g2o::RawLaser::Point2DVector& out);

}  // namespace cgm

class ScanMatcher {
 public:
  ScanMatcher() : _kernelRange(0.0), _kernelResolution(0.0), _kscale(128), _device(0) {}
  void setDevice(int device) { _device = device; }

  // scan_matcher.cpp:38-61. The stamp itself is built inside the library from these two numbers
  // when the grid is created, so the call order of the reference (kernel first, then grid;
  // graph_slam.cpp:59-62) is kept.
  void initializeKernel(double resolution, double kernelRange) {
    _kernelResolution = resolution;
    _kernelRange = kernelRange;
  }
  void initializeGrid(Eigen::Vector2f lowerLeft, Eigen::Vector2f upperRight, double resolution) {
    _grid = CharGrid(lowerLeft, upperRight, static_cast<float>(resolution), _kscale,
                     _kernelResolution > 0.0 ? _kernelResolution : resolution, _kernelRange,
                     _device);
  }
  void resetGrid() { _grid.reset(); }  // scan_matcher.cpp:68-76

  static void applyTransfToScan(g2o::SE2 transf, const g2o::RawLaser::Point2DVector& scan,
                                g2o::RawLaser::Point2DVector& outScan) {  // :78-87
    outScan.resize(scan.size());
    for (size_t i = 0; i < scan.size(); ++i) {
      g2o::SE2 point;
      point.setTranslation(scan[i]);
      outScan[i] = (transf * point).translation();
    }
  }

  bool closeScanMatching(g2o::OptimizableGraph::VertexSet& vset,
                         g2o::OptimizableGraph::Vertex* _originVertex,
                         g2o::OptimizableGraph::Vertex* _currentVertex, g2o::SE2* trel,
                         double maxScore) {  // scan_matcher.cpp:112-189
    g2o::VertexSE2* currentVertex = dynamic_cast<g2o::VertexSE2*>(_currentVertex);
    g2o::VertexSE2* originVertex = dynamic_cast<g2o::VertexSE2*>(_originVertex);
    if (!currentVertex || !originVertex) return false;
    resetGrid();
    g2o::RawLaser::Point2DVector scansInRefVertex;
    cgm::transformPointsFromVSet(vset, _originVertex, scansInRefVertex);
    _grid.addAndConvolvePoints<g2o::RawLaser::Point2DVector>(scansInRefVertex.begin(),
                                                             scansInRefVertex.end());
    g2o::RobotLaser* lasercv = dynamic_cast<g2o::RobotLaser*>(currentVertex->userData());
    if (!lasercv) return false;
    g2o::RawLaser::Point2DVector cvscan = lasercv->cartesian();
    Vector2dVector reducedCvscan;
    CharGrid::subsample(reducedCvscan, cvscan, 0.1);
    g2o::RawLaser::Point2DVector cvScanRobot;
    applyTransfToScan(lasercv->laserParams().laserPose, reducedCvscan, cvScanRobot);

    const g2o::SE2 delta = originVertex->estimate().inverse() * currentVertex->estimate();
    const Eigen::Vector3d initGuess(delta.translation().x(), delta.translation().y(),
                                    delta.rotation().angle());
    std::vector<MatcherResult> mresvec;
    const double thetaRes = 0.0125 * .5;
    // double sums narrowed to float bounds, as the reference's Vector3f construction does
    Eigen::Vector3f lower(-.3 + initGuess.x(), -.3 + initGuess.y(), -0.2 + initGuess.z());
    Eigen::Vector3f upper(+.3 + initGuess.x(), .3 + initGuess.y(), 0.2 + initGuess.z());
    const clock_t t_ini = clock();
    _grid.greedySearch(mresvec, cvScanRobot, lower, upper, thetaRes, maxScore, 0.5, 0.5, 0.2);
    const double secs = static_cast<double>(clock() - t_ini) / CLOCKS_PER_SEC;
    printf("Greedy search: %.16g ms. Matcher results: %i\n", secs * 1000.0,
           static_cast<int>(mresvec.size()));
    if (mresvec.size()) {
      const Eigen::Vector3d adj = mresvec[0].transformation;
      trel->setTranslation(Eigen::Vector2d(adj.x(), adj.y()));
      trel->setRotation(Eigen::Rotation2Dd(adj.z()));
      return true;
    }
    std::cerr << std::endl;
    return false;
  }

  bool scanMatchingLC(g2o::OptimizableGraph::VertexSet& referenceVset,
                      g2o::OptimizableGraph::Vertex* _referenceVertex,
                      g2o::OptimizableGraph::Vertex* _currentVertex, std::vector<g2o::SE2>& trel,
                      double maxScore) {  // :191-199
    g2o::OptimizableGraph::VertexSet currvset;
    currvset.insert(_currentVertex);
    return scanMatchingLC(referenceVset, _referenceVertex, currvset, _currentVertex, trel, maxScore);
  }

  bool scanMatchingLC(g2o::OptimizableGraph::VertexSet& referenceVset,
                      g2o::OptimizableGraph::Vertex* _referenceVertex,
                      g2o::OptimizableGraph::VertexSet& currvset,
                      g2o::OptimizableGraph::Vertex* _currentVertex, std::vector<g2o::SE2>& trel,
                      double maxScore) {  // :201-294
    std::cerr << "Loop Closing Scan Matching" << std::endl;
    g2o::VertexSE2* referenceVertex = dynamic_cast<g2o::VertexSE2*>(_referenceVertex);
    if (!referenceVertex) return false;
    resetGrid();
    trel.clear();
    g2o::RawLaser::Point2DVector scansInRefVertex;
    cgm::transformPointsFromVSet(referenceVset, _referenceVertex, scansInRefVertex);
    _grid.addAndConvolvePoints<g2o::RawLaser::Point2DVector>(scansInRefVertex.begin(),
                                                             scansInRefVertex.end());
    g2o::RawLaser::Point2DVector scansInCurVertex;
    cgm::transformPointsFromVSet(currvset, _currentVertex, scansInCurVertex);
    Vector2dVector reducedScans;
    CharGrid::subsample(reducedScans, scansInCurVertex, 0.1);

    RegionVector regions, regionspi;
    for (g2o::OptimizableGraph::VertexSet::iterator it = referenceVset.begin();
         it != referenceVset.end(); ++it) {
      g2o::VertexSE2* vertex = static_cast<g2o::VertexSE2*>(*it);
      g2o::SE2 relposv(.0, .0, .0);
      if (vertex->id() != referenceVertex->id())
        relposv = referenceVertex->estimate().inverse() * vertex->estimate();
      Eigen::Vector3f lower(-.5 + relposv.translation().x(), -1.5 + relposv.translation().y(),
                            -0.8 + relposv.rotation().angle());
      Eigen::Vector3f upper(.5 + relposv.translation().x(), 1.5 + relposv.translation().y(),
                            0.8 + relposv.rotation().angle());
      Region reg;
      reg.lowerLeft = lower;
      reg.upperRight = upper;
      regions.push_back(reg);
      lower[2] += M_PI;  // float += double, rounded to float (scan_matcher.cpp:236-237)
      upper[2] += M_PI;
      reg.lowerLeft = lower;
      reg.upperRight = upper;
      regionspi.push_back(reg);
    }
    std::vector<MatcherResult> mresvec;
    const double thetaRes = 0.025, dx = 0.5, dy = 0.5, dth = 0.2;
    std::map<DiscreteTriplet, MatcherResult> resultsMap;
    for (int pass = 0; pass < 2; ++pass) {
      const clock_t t_ini = clock();
      _grid.greedySearch(mresvec, reducedScans, pass ? regionspi : regions, thetaRes, maxScore, dx,
                         dy, dth);
      const double secs = static_cast<double>(clock() - t_ini) / CLOCKS_PER_SEC;
      printf("%.16g ms. Matcher results: %i\n", secs * 1000.0, static_cast<int>(mresvec.size()));
      if (mresvec.size()) {
        mresvec[0].transformation[2] = g2o::normalize_theta(mresvec[0].transformation[2]);
        std::cerr << (pass ? "Found Loop Closure Edge PI. Transf: " : "Found Loop Closure Edge. Transf: ")
                  << mresvec[0].transformation.x() << " " << mresvec[0].transformation.y() << " "
                  << mresvec[0].transformation.z() << std::endl;
        CharGrid::addToPrunedMap(resultsMap, mresvec[0], dx, dy, dth);
      }
    }
    for (std::map<DiscreteTriplet, MatcherResult>::iterator it = resultsMap.begin();
         it != resultsMap.end(); ++it) {
      const Eigen::Vector3d adj = it->second.transformation;
      g2o::SE2 transf;
      transf.setTranslation(Eigen::Vector2d(adj.x(), adj.y()));
      transf.setRotation(Eigen::Rotation2Dd(adj.z()));
      trel.push_back(transf);
      std::cerr << "Final result: " << transf.translation().x() << " " << transf.translation().y()
                << " " << transf.rotation().angle() << std::endl;
    }
    return !trel.empty();
  }

  bool scanMatchingLChierarchical(g2o::OptimizableGraph::VertexSet& referenceVset,
                                  g2o::OptimizableGraph::Vertex* _referenceVertex,
                                  g2o::OptimizableGraph::VertexSet& currvset,
                                  g2o::OptimizableGraph::Vertex* _currentVertex,
                                  std::vector<g2o::SE2>& trel, double maxScore) {  // :296-355
    g2o::VertexSE2* currentVertex = dynamic_cast<g2o::VertexSE2*>(_currentVertex);
    g2o::VertexSE2* referenceVertex = dynamic_cast<g2o::VertexSE2*>(_referenceVertex);
    if (!currentVertex || !referenceVertex) return false;
    resetGrid();
    trel.clear();
    g2o::RawLaser::Point2DVector scansInRefVertex;
    cgm::transformPointsFromVSet(referenceVset, _referenceVertex, scansInRefVertex);
    _grid.addAndConvolvePoints<g2o::RawLaser::Point2DVector>(scansInRefVertex.begin(),
                                                             scansInRefVertex.end());
    g2o::RawLaser::Point2DVector scansInCurVertex;
    cgm::transformPointsFromVSet(currvset, _currentVertex, scansInCurVertex);
    Vector2dVector reducedScans;
    CharGrid::subsample(reducedScans, scansInCurVertex, 0.1);
    const g2o::SE2 delta = referenceVertex->estimate().inverse() * currentVertex->estimate();
    const Eigen::Vector3d g(delta.translation().x(), delta.translation().y(), delta.rotation().angle());
    RegionVector regions(1);
    regions[0].lowerLeft = Eigen::Vector3f(-2. + g.x(), -2. + g.y(), -1. + g.z());
    regions[0].upperRight = Eigen::Vector3f(+2. + g.x(), 2. + g.y(), 1. + g.z());
    std::vector<MatcherResult> mresvec;
    _grid.hierarchicalSearch(mresvec, reducedScans, regions, 0.025, maxScore, 0.5, 0.5, 0.2, 3);
    if (mresvec.size()) {
      const Eigen::Vector3d adj = mresvec[0].transformation;
      g2o::SE2 transf;
      transf.setTranslation(Eigen::Vector2d(adj.x(), adj.y()));
      transf.setRotation(Eigen::Rotation2Dd(adj.z()));
      trel.push_back(transf);
    }
    return !trel.empty();
  }

  bool globalMatching(g2o::OptimizableGraph::VertexSet& referenceVset,
                      g2o::OptimizableGraph::Vertex* _referenceVertex,
                      g2o::OptimizableGraph::Vertex* _currentVertex, g2o::SE2* trel,
                      double maxScore) {  // :357-364
    g2o::OptimizableGraph::VertexSet vset;
    vset.insert(_currentVertex);
    return globalMatching(referenceVset, _referenceVertex, vset, _currentVertex, trel, maxScore);
  }

  bool globalMatching(g2o::OptimizableGraph::VertexSet& referenceVset,
                      g2o::OptimizableGraph::Vertex* _referenceVertex,
                      g2o::OptimizableGraph::VertexSet& currvset,
                      g2o::OptimizableGraph::Vertex* _currentVertex, g2o::SE2* trel,
                      double maxScore) {  // :366-428
    resetGrid();
    g2o::RawLaser::Point2DVector scansInRefVertex;
    cgm::transformPointsFromVSet(referenceVset, _referenceVertex, scansInRefVertex);
    _grid.addAndConvolvePoints<g2o::RawLaser::Point2DVector>(scansInRefVertex.begin(),
                                                             scansInRefVertex.end());
    g2o::RawLaser::Point2DVector scansInCurVertex;
    cgm::transformPointsFromVSet(currvset, _currentVertex, scansInCurVertex);
    std::vector<MatcherResult> mresvec;
    Eigen::Vector3f lower(-10, -5, -M_PI);
    Eigen::Vector3f upper(+10, +5, M_PI);
    Vector2dVector reducedScans;
    CharGrid::subsample(reducedScans, scansInCurVertex, 0.1);
    _grid.hierarchicalSearch(mresvec, reducedScans, lower, upper, 0.025, maxScore, 0.5, 0.5, 0.2, 4);
    if (mresvec.size()) {
      const Eigen::Vector3d adj = mresvec[0].transformation;
      trel->setTranslation(Eigen::Vector2d(adj.x(), adj.y()));
      trel->setRotation(Eigen::Rotation2Dd(adj.z()));
      return true;
    }
    return false;
  }

  bool verifyMatching(g2o::OptimizableGraph::VertexSet& vset1,
                      g2o::OptimizableGraph::Vertex* _referenceVertex1,
                      g2o::OptimizableGraph::VertexSet& vset2,
                      g2o::OptimizableGraph::Vertex* _referenceVertex2, g2o::SE2 trel12,
                      double* score) {  // :430-505
    g2o::VertexSE2* referenceVertex2 = dynamic_cast<g2o::VertexSE2*>(_referenceVertex2);
    if (!referenceVertex2) return false;
    resetGrid();
    CharGrid auxGrid = _grid.clone();  // `CharGrid auxGrid = _grid;` copies the cells
    g2o::RawLaser::Point2DVector scansvset2inref1;
    for (g2o::OptimizableGraph::VertexSet::iterator it = vset2.begin(); it != vset2.end(); ++it) {
      g2o::VertexSE2* vertex = static_cast<g2o::VertexSE2*>(*it);
      g2o::RobotLaser* laserv = dynamic_cast<g2o::RobotLaser*>(vertex->userData());
      if (!laserv) continue;
      g2o::RawLaser::Point2DVector vscan = laserv->cartesian(), in_ref1;
      const g2o::SE2 trl = laserv->laserParams().laserPose;
      if (vertex->id() == referenceVertex2->id()) {
        applyTransfToScan(trel12 * trl, vscan, in_ref1);
      } else {
        const g2o::SE2 tref2_v = referenceVertex2->estimate().inverse() * vertex->estimate();
        applyTransfToScan(trel12 * tref2_v * trl, vscan, in_ref1);
      }
      scansvset2inref1.insert(scansvset2inref1.end(), in_ref1.begin(), in_ref1.end());
    }
    g2o::RawLaser::Point2DVector scansvset1;
    cgm::transformPointsFromVSet(vset1, _referenceVertex1, scansvset1);
    _grid.addAndConvolvePoints<g2o::RawLaser::Point2DVector>(scansvset2inref1.begin(),
                                                             scansvset2inref1.end());
    g2o::RawLaser::Point2DVector nonmatchedpoints;
    _grid.searchNonMatchedPoints(scansvset1, nonmatchedpoints, .3);
    auxGrid.addAndConvolvePoints<g2o::RawLaser::Point2DVector>(nonmatchedpoints.begin(),
                                                              nonmatchedpoints.end());
    Eigen::Vector2f lower(-.3 + trel12.translation().x(), -.3 + trel12.translation().y());
    Eigen::Vector2f upper(+.3 + trel12.translation().x(), +.3 + trel12.translation().y());
    auxGrid.countPoints(lower, upper, score);
    std::cerr << "Score: " << *score << std::endl;
    return *score <= 40.0;
  }

  // The reference returns a copy of the grid; here the copy shares the device handle
  // (read access through grid().grid().cell()).
  inline CharGrid grid() const { return _grid; }

 protected:
  CharGrid _grid;
  double _kernelRange, _kernelResolution;
  int _kscale, _device;
};

namespace cgm {

inline void transformPointsFromVSet(g2o::OptimizableGraph::VertexSet& vset,
                                    g2o::OptimizableGraph::Vertex* reference_,
                                    g2o::RawLaser::Point2DVector& out) {
  g2o::VertexSE2* reference = dynamic_cast<g2o::VertexSE2*>(reference_);
  out.clear();
  if (!reference) return;
  for (g2o::OptimizableGraph::VertexSet::iterator it = vset.begin(); it != vset.end(); ++it) {
    g2o::VertexSE2* vertex = static_cast<g2o::VertexSE2*>(*it);
    g2o::RobotLaser* laserv = dynamic_cast<g2o::RobotLaser*>(vertex->userData());
    if (!laserv) continue;
    const g2o::RawLaser::Point2DVector vscan = laserv->cartesian();
    const g2o::SE2 trl = laserv->laserParams().laserPose;
    g2o::RawLaser::Point2DVector in_ref;
    if (vertex->id() == reference->id()) {
      ScanMatcher::applyTransfToScan(trl, vscan, in_ref);
    } else {
      const g2o::SE2 trel = reference->estimate().inverse() * vertex->estimate();
      ScanMatcher::applyTransfToScan(trel * trl, vscan, in_ref);
    }
    out.insert(out.end(), in_ref.begin(), in_ref.end());
  }
}

}  // namespace cgm

#endif
