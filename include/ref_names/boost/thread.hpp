// Stand-in for <boost/thread.hpp> (Boost is not in this image): the reference's SLAM classes use
// boost::mutex + boost::mutex::scoped_lock (src/slam/graph_slam.h:119, graph_slam.cpp:48 ...) and
// GraphComm uses boost::thread (src/mrslam/graph_comm.cpp:56-58). With Boost installed, drop this
// directory from the include path and the real header is used.
#ifndef CGM_REF_NAMES_BOOST_THREAD_HPP
#define CGM_REF_NAMES_BOOST_THREAD_HPP
#include <mutex>
#include <thread>

namespace boost {
class mutex {
 public:
  class scoped_lock {
   public:
    explicit scoped_lock(mutex& m) : l_(m.m_) {}

   private:
    std::unique_lock<std::mutex> l_;
  };

 private:
  std::mutex m_;
};
typedef std::thread thread;
}  // namespace boost
#endif
