// Scenario driver over the REFERENCE'S OWN host classes -- ScanMatcher, VerticesFinder,
// LoopClosureChecker, ClosureBuffer, MRClosureBuffer, CondensedGraphBuffer / CondensedGraphCreator,
// GraphSLAM::checkCovariance / addNeighboringVertices -- compiled verbatim from /root/reference
// (oracle/Makefile, target `frontend`) against include/g2o_compat + include/cgm/chargrid.hpp
// (compat_driver_gpu, linked to libcgmrslam_b200.so) or against the reference's chargrid.cpp and the
// CPU oracle solver (compat_driver_cpu). Reads a scenario file, prints results with 17 significant
// digits; tests/test_cpp_compat.py and tests/test_frontend_host.py compare them with the CPU
// oracles, which pins those restatements to the reference's code.
//
//   V id x y th fixed n_beams first_angle step max_range r_1 ... r_n     vertex + its scan
//   E i j dx dy dth I11 I12 I13 I22 I23 I33                              EdgeSE2
//   OPT n                                            initializeOptimization(); optimize(n)
//   POSES                                            print all estimates
//   CLOSE origin current max_score k id_1 ... id_k   closeScanMatching on the close matcher
//   LC reference current max_score k id_1 ... id_k   scanMatchingLC on the LC matcher
//   GLOBAL reference current max_score k id_1..id_k  globalMatching on the LC matcher
//   MARG k id_1 ... id_k                             computeMarginals diagonal blocks
//   STAR gauge k id_1 ... id_k                       CondensedGraphCreator::compute pattern
//   SAVE path / LOAD path                            g2o text round trip
//   FINDSM cur                                       VerticesFinder::findVerticesScanMatching
//   SETS cur                                         ... + findSetsOfVertices + findClosestVertex
//   NEIGH cur gap k id_1 ... id_k                    addNeighboringVertices
//   COVGATE cur k id_1 ... id_k                      checkCovariance (marginals on the GPU)
//   CLOSURES thr nv id_1..id_nv ne (from to dx dy dth)*ne   LoopClosureChecker::init + check
//   BUFSIM window n (vertex_id n_edges)*n            ClosureBuffer add / updateList / checkList
//   MRBUF window n (op robot vertex_id n_edges)*n    MRClosureBuffer insert (I) / remove (R) / update (U)
//   CGB robot optimal k id_1 ... id_k                CondensedGraphBuffer::insertOutClosure + computeCondensedGraph
//   CGIN robot n (from to dx dy dth I11..I33)*n      CondensedGraphBuffer::insertEdgesFromRobot
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "mrslam/condensed_graph/condensed_graph_buffer.h"
#include "mrslam/mr_closure_buffer.h"
#include "slam/graph_slam.h"

#include "driver_out.h"

using namespace g2o;

// GraphSLAM keeps its candidate filters protected; a subclass may call them (no source edit)
struct Probe : public GraphSLAM {
  void neighbours(VertexSE2* last, OptimizableGraph::VertexSet& vset, int gap) {
    _lastVertex = last;
    addNeighboringVertices(vset, gap);
  }
  void covarianceGate(VertexSE2* last, OptimizableGraph::VertexSet& vset) {
    _lastVertex = last;
    for (HyperGraph::VertexIDMap::iterator it = _graph->vertices().begin(); it != _graph->vertices().end(); ++it)
      if (static_cast<OptimizableGraph::Vertex*>(it->second)->fixed()) _firstRobotPose = static_cast<VertexSE2*>(it->second);
    checkCovariance(vset);
  }
};

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream f(argv[1]);
  if (!f) return 2;
  Probe probe;
  SparseOptimizer& opt = *probe.graph();
  // the reference's construction idiom (graph_slam.cpp:44-55) must keep compiling
  typedef BlockSolver<BlockSolverTraits<-1, -1> > SlamBlockSolver;
  typedef LinearSolverCSparse<SlamBlockSolver::PoseMatrixType> SlamLinearSolver;
  auto linearSolver = std::unique_ptr<SlamLinearSolver>(new SlamLinearSolver());
  linearSolver->setBlockOrdering(false);
  auto blockSolver = std::unique_ptr<SlamBlockSolver>(new SlamBlockSolver(std::move(linearSolver)));
  opt.setAlgorithm(new OptimizationAlgorithmGaussNewton(std::move(blockSolver)));
  opt.setVerbose(false);

  CondensedGraphBuffer condensedGraphs(&opt);   // as MRGraphSLAM holds one (mr_graph_slam.cpp:32)
  ScanMatcher closeMatcher, lcMatcher;
  bool matchers_ready = false;
  auto ready = [&]() {
    if (matchers_ready) return;
    closeMatcher.initializeKernel(0.025, 0.2);  // graph_slam.cpp:59-62
    closeMatcher.initializeGrid(Eigen::Vector2f(-15, -15), Eigen::Vector2f(15, 15), 0.025);
    lcMatcher.initializeKernel(0.1, 0.5);
    lcMatcher.initializeGrid(Eigen::Vector2f(-35, -35), Eigen::Vector2f(35, 35), 0.1);
    matchers_ready = true;
  };
  std::string line, tag;
  printf("BEGIN\n");
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    if (!(ss >> tag)) continue;
    if (tag == "V") {
      int id, fixed, nb;
      double x, y, th, first, step, maxr;
      ss >> id >> x >> y >> th >> fixed >> nb >> first >> step >> maxr;
      VertexSE2* v = new VertexSE2();
      v->setId(id);
      v->setEstimate(SE2(x, y, th));
      v->setFixed(fixed != 0);
      if (nb > 0) {
        std::vector<double> r(nb);
        for (int i = 0; i < nb; ++i) ss >> r[i];
        RobotLaser* rl = new RobotLaser();
        LaserParameters lp(0, nb, first, step, maxr, 0.1, 0);
        lp.laserPose = SE2(0.05, 0.0, 0.0);
        rl->setLaserParams(lp);
        rl->setRanges(r);
        v->setUserData(rl);
      }
      opt.addVertex(v);
    } else if (tag == "E") {
      int a, b;
      double x, y, th, w[6];
      ss >> a >> b >> x >> y >> th;
      for (int k = 0; k < 6; ++k) ss >> w[k];
      EdgeSE2* e = new EdgeSE2();
      e->vertices()[0] = opt.vertex(a);
      e->vertices()[1] = opt.vertex(b);
      e->setMeasurement(SE2(x, y, th));
      Eigen::Matrix3d m;
      m(0, 0) = w[0]; m(0, 1) = m(1, 0) = w[1]; m(0, 2) = m(2, 0) = w[2];
      m(1, 1) = w[3]; m(1, 2) = m(2, 1) = w[4]; m(2, 2) = w[5];
      e->setInformation(m);
      opt.addEdge(e);
    } else if (tag == "OPT") {
      int n;
      ss >> n;
      opt.initializeOptimization();
      printf("OPT %d\n", opt.optimize(n));
    } else if (tag == "POSES") {
      for (auto& kv : opt.vertices()) {
        const VertexSE2* v = static_cast<const VertexSE2*>(kv.second);
        printf("P %d %.17g %.17g %.17g\n", v->id(), v->estimate().translation().x(),
               v->estimate().translation().y(), v->estimate().rotation().angle());
      }
    } else if (tag == "CLOSE" || tag == "LC" || tag == "GLOBAL") {
      ready();
      int ref, cur, k;
      double max_score;
      ss >> ref >> cur >> max_score >> k;
      OptimizableGraph::VertexSet vset;
      for (int i = 0; i < k; ++i) {
        int id;
        ss >> id;
        vset.insert(opt.vertex(id));
      }
      if (tag == "CLOSE") {
        SE2 t;
        const bool ok = closeMatcher.closeScanMatching(vset, opt.vertex(ref), opt.vertex(cur), &t, max_score);
        printf("CLOSE %d %.17g %.17g %.17g\n", ok ? 1 : 0, t.translation().x(), t.translation().y(),
               t.rotation().angle());
      } else if (tag == "LC") {
        std::vector<SE2> ts;
        const bool ok = lcMatcher.scanMatchingLC(vset, opt.vertex(ref), opt.vertex(cur), ts, max_score);
        printf("LC %d %d\n", ok ? 1 : 0, static_cast<int>(ts.size()));
        for (auto& t : ts)
          printf("T %.17g %.17g %.17g\n", t.translation().x(), t.translation().y(), t.rotation().angle());
      } else {
        SE2 t;
        const bool ok = lcMatcher.globalMatching(vset, opt.vertex(ref), opt.vertex(cur), &t, max_score);
        printf("GLOBAL %d %.17g %.17g %.17g\n", ok ? 1 : 0, t.translation().x(), t.translation().y(),
               t.rotation().angle());
      }
    } else if (tag == "MARG") {
      int k;
      ss >> k;
      std::vector<std::pair<int, int> > idx;
      for (int i = 0; i < k; ++i) {
        int id;
        ss >> id;
        const int h = opt.vertex(id)->hessianIndex();
        idx.push_back(std::make_pair(h, h));
      }
      SparseBlockMatrix<Eigen::MatrixXd> spinv;
      const bool ok = opt.computeMarginals(spinv, idx);
      printf("MARG %d\n", ok ? 1 : 0);
      for (auto& p : idx) {
        Eigen::MatrixXd* b = spinv.block(p.first, p.second);
        printf("M");
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) printf(" %.17g", b ? (*b)(i, j) : 0.0);
        printf("\n");
      }
    } else if (tag == "STAR") {
      // GraphManipulator::pushState/fixGauge/optimize + CondensedGraphCreator::compute
      // (graph_manipulator.cpp:62-124, condensed_graph_creator.cpp:33-66)
      int gauge, k;
      ss >> gauge >> k;
      std::vector<int> ids(k);
      for (int i = 0; i < k; ++i) ss >> ids[i];
      std::vector<bool> was_fixed;
      for (auto& kv : opt.vertices()) {
        OptimizableGraph::Vertex* v = static_cast<OptimizableGraph::Vertex*>(kv.second);
        v->push();
        was_fixed.push_back(v->fixed());
        v->setFixed(v->id() == gauge);
      }
      HyperGraph::EdgeSet es = opt.edges();
      opt.initializeOptimization(es);
      opt.computeInitialGuess();
      const int done = opt.optimize(1);
      std::set<OptimizableGraph::Edge*> star;
      std::vector<EdgeSE2*> order;
      for (int id : ids) {
        if (id == gauge) continue;
        EdgeSE2* e = new EdgeSE2();
        e->vertices()[0] = opt.vertex(gauge);
        e->vertices()[1] = opt.vertex(id);
        star.insert(e);
        order.push_back(e);
      }
      EdgeLabeler labeler(&opt);
      const int labelled = labeler.labelEdges(star);
      printf("STAR %d %d\n", done, labelled);
      for (EdgeSE2* e : order) {
        printf("S %d %.17g %.17g %.17g", e->vertex(1)->id(), e->measurement().translation().x(),
               e->measurement().translation().y(), e->measurement().rotation().angle());
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) printf(" %.17g", e->information()(i, j));
        printf("\n");
        delete e;
      }
      size_t q = 0;
      for (auto& kv : opt.vertices()) {
        OptimizableGraph::Vertex* v = static_cast<OptimizableGraph::Vertex*>(kv.second);
        v->pop();
        v->setFixed(was_fixed[q++]);
      }
    } else if (tag == "MRBUF") {
      // the per-peer windows as MRGraphSLAM drives them (mr_graph_slam.cpp:233-247,310-327)
      int window, n;
      ss >> window >> n;
      MRClosureBuffer mr;
      std::vector<EdgeSE2*> owned;
      std::map<int, std::vector<EdgeSE2*> > edges_of;   // by vertex id
      for (int step = 0; step < n; ++step) {
        std::string op;
        int robot, vid, ne;
        ss >> op >> robot >> vid >> ne;
        if (op == "I") {
          ClosureBuffer c;
          c.addVertex(opt.vertex(vid));
          for (int i = 0; i < ne; ++i) {
            EdgeSE2* e = new EdgeSE2();
            e->vertices()[0] = opt.vertices().begin()->second;
            e->vertices()[1] = opt.vertex(vid);
            owned.push_back(e);
            edges_of[vid].push_back(e);
            c.addEdge(e);
          }
          mr.insert(c, robot);
        } else if (op == "R") {
          ClosureBuffer c;
          c.addVertex(opt.vertex(vid));
          for (EdgeSE2* e : edges_of[vid]) c.addEdge(e);
          mr.remove(c, robot);
        } else {
          mr.update(window);
        }
        MRClosureBuffer snapshot = mr;   // findInterRobotConstraints iterates over a copy (mr_graph_slam.cpp:270)
        printf("MR %u", snapshot.size());
        for (auto& kv : snapshot.mrClosures) {
          printf(" | %d %zu %zu", kv.first, kv.second->edgeSet().size(), kv.second->vertices().size());
          for (const VertexTime& vt : kv.second->vertexList()) printf(" %d:%d", vt.v->id(), vt.time);
        }
        printf("\n");
      }
      for (EdgeSE2* e : owned) delete e;
    } else if (tag == "CGB") {
      // the star another robot gets over the vertices it asked about (condensed_graph_buffer.cpp:437-485)
      int robot, optimal, k;
      ss >> robot >> optimal >> k;
      OptimizableGraph::VertexIDMap asked;
      for (int i = 0; i < k; ++i) {
        int id;
        ss >> id;
        asked.insert(std::make_pair(id, opt.vertex(id)));
      }
      condensedGraphs.insertOutClosure(robot, asked);
      condensedGraphs.computeCondensedGraph(robot, optimal != 0);
      OptimizableGraph::EdgeSet star = condensedGraphs.outCondensedGraph(robot);
      const int gauge = star.empty() ? -1 : (*star.begin())->vertex(0)->id();
      int at_level = 0;
      for (HyperGraph::Edge* he : opt.edges())
        at_level += static_cast<OptimizableGraph::Edge*>(he)->level() == robot + 1;
      printf("CGB %d %zu %zu %zu %d %zu\n", gauge, star.size(), opt.edges().size(),
             condensedGraphs.getMyEdges().size(), at_level, condensedGraphs.outClosures(robot)->size());
      for (HyperGraph::Edge* he : star) {
        EdgeSE2* e = static_cast<EdgeSE2*>(he);
        printf("C %d %.17g %.17g %.17g", e->vertex(1)->id(), e->measurement().translation().x(),
               e->measurement().translation().y(), e->measurement().rotation().angle());
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) printf(" %.17g", e->information()(i, j));
        printf("\n");
      }
    } else if (tag == "CGIN") {
      // a star received from another robot replaces the previous one (condensed_graph_buffer.cpp:487-510)
      int robot, n;
      ss >> robot >> n;
      OptimizableGraph::EdgeSet eset;
      for (int i = 0; i < n; ++i) {
        int a, b;
        double x, y, th, w[6];
        ss >> a >> b >> x >> y >> th;
        for (int q = 0; q < 6; ++q) ss >> w[q];
        EdgeSE2* e = new EdgeSE2();
        e->vertices()[0] = opt.vertex(a);
        e->vertices()[1] = opt.vertex(b);
        e->setMeasurement(SE2(x, y, th));
        Eigen::Matrix3d m;
        m(0, 0) = w[0]; m(0, 1) = m(1, 0) = w[1]; m(0, 2) = m(2, 0) = w[2];
        m(1, 1) = w[3]; m(1, 2) = m(2, 1) = w[4]; m(2, 2) = w[5];
        e->setInformation(m);
        eset.insert(e);
      }
      condensedGraphs.insertEdgesFromRobot(robot, eset);
      printf("CGIN %zu %zu %zu\n", condensedGraphs.inCondensedGraph(robot).size(), opt.edges().size(),
             condensedGraphs.getMyEdges().size());
    } else if (tag == "FINDSM" || tag == "SETS") {
      int cur;
      ss >> cur;
      VerticesFinder vf(&opt);
      OptimizableGraph::VertexSet vset;
      vf.findVerticesScanMatching(opt.vertex(cur), vset);
      if (tag == "FINDSM") {
        printf("FINDSM %zu", vset.size());
        for (HyperGraph::Vertex* v : vset) printf(" %d", v->id());
        printf("\n");
      } else {
        std::set<OptimizableGraph::VertexSet> sets;
        vf.findSetsOfVertices(vset, sets);
        printf("SETS %zu\n", sets.size());
        for (OptimizableGraph::VertexSet g : sets) {
          printf("G %d %zu", vf.findClosestVertex(g, opt.vertex(cur))->id(), g.size());
          for (HyperGraph::Vertex* v : g) printf(" %d", v->id());
          printf("\n");
        }
      }
    } else if (tag == "NEIGH" || tag == "COVGATE") {
      int cur, gap = 0, k;
      ss >> cur;
      if (tag == "NEIGH") ss >> gap;
      ss >> k;
      OptimizableGraph::VertexSet vset;
      for (int i = 0; i < k; ++i) {
        int id;
        ss >> id;
        vset.insert(opt.vertex(id));
      }
      VertexSE2* last = static_cast<VertexSE2*>(opt.vertex(cur));
      if (tag == "NEIGH") probe.neighbours(last, vset, gap);
      else probe.covarianceGate(last, vset);
      printf("%s %zu", tag.c_str(), vset.size());
      for (HyperGraph::Vertex* v : vset) printf(" %d", v->id());
      printf("\n");
    } else if (tag == "CLOSURES") {
      double thr;
      int nv, ne;
      ss >> thr >> nv;
      OptimizableGraph::VertexIDMap window;
      for (int i = 0; i < nv; ++i) {
        int id;
        ss >> id;
        window[id] = opt.vertex(id);
      }
      ss >> ne;
      OptimizableGraph::EdgeSet cand;
      std::vector<EdgeSE2*> order;
      for (int i = 0; i < ne; ++i) {
        int a, b;
        double dx, dy, dth;
        ss >> a >> b >> dx >> dy >> dth;
        EdgeSE2* e = new EdgeSE2();
        e->vertices()[0] = opt.vertex(a);
        e->vertices()[1] = opt.vertex(b);
        e->setMeasurement(SE2(dx, dy, dth));
        Eigen::Matrix3d info = Eigen::Matrix3d::Identity();
        info(0, 0) = info(1, 1) = 1000.0;
        info(2, 2) = 10000.0;  // _SMinf, graph_slam.cpp:75-76
        e->setInformation(info);
        cand.insert(e);
        order.push_back(e);
      }
      LoopClosureChecker lcc;
      lcc.init(window, cand, thr);
      lcc.check();
      printf("CLOSURES %d %.17g\n", lcc.inliers(), lcc.chi2());
      for (EdgeSE2* e : order) printf("C %.17g\n", lcc.closures()[e]);
      for (EdgeSE2* e : order) delete e;
    } else if (tag == "BUFSIM") {
      int window, n;
      ss >> window >> n;
      ClosureBuffer buf;
      std::vector<EdgeSE2*> owned;
      for (int step = 0; step < n; ++step) {
        int vid, ne;
        ss >> vid >> ne;
        if (ne > 0) {  // GraphSLAM::addClosures: the candidates of this keyframe + the keyframe
          OptimizableGraph::EdgeSet es;
          for (int i = 0; i < ne; ++i) {
            EdgeSE2* e = new EdgeSE2();
            e->vertices()[0] = opt.vertices().begin()->second;
            e->vertices()[1] = opt.vertex(vid);
            owned.push_back(e);
            es.insert(e);
          }
          buf.addEdgeSet(es);
          buf.addVertex(opt.vertex(vid));
        }
        const bool vote = buf.checkList(window);  // GraphSLAM::checkClosures, then updateClosures
        buf.updateList(window);
        printf("B %d %zu %zu", vote ? 1 : 0, buf.edgeSet().size(), buf.vertices().size());
        for (const VertexTime& vt : buf.vertexList()) printf(" %d:%d", vt.v->id(), vt.time);
        printf("\n");
      }
      for (EdgeSE2* e : owned) delete e;
    } else if (tag == "SAVE") {
      std::string path;
      ss >> path;
      printf("SAVE %d\n", opt.save(path.c_str()) ? 1 : 0);
    }
  }
  printf("END\n");
  return 0;
}
