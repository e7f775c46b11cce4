// TEST ONLY. 200 000 truncated, bit-flipped and random datagrams through MessageFactory::fromCharArray
// (include/cgm/msg_factory.hpp), each in a heap buffer of exactly its size, built with
// -fsanitize=address,undefined by tests/test_msg_wire.py: any read past the datagram or any
// undefined behaviour aborts the run.
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include "cgm/msg_factory.hpp"
int main() {
  std::mt19937 rng(7);
  MessageFactory f;
  f.registerMessageType<VertexArrayMessage>(); f.registerMessageType<RobotLaserMessage>();
  f.registerMessageType<ComboMessage>(); f.registerMessageType<EdgeArrayMessage>();
  f.registerMessageType<ClosuresMessage>(); f.registerMessageType<CondensedGraphMessage>();
  f.registerMessageType<GraphMessage>();
  GraphMessage g(3);
  g.vertexVector.resize(7); g.edgeVector.resize(5); g.closures.assign(9, 42);
  ComboMessage c(1); c.vertexVector.resize(5); c.readings.assign(361, 1.5); c.nodeId = 10007;
  std::vector<char> base(100000);
  std::vector<std::vector<char>> valid;
  for (RobotMessage* m : {static_cast<RobotMessage*>(&g), static_cast<RobotMessage*>(&c)}) {
    char* e = m->toCharArray(base.data(), base.size());
    valid.emplace_back(base.data(), e);
  }
  long ok = 0, bad = 0;
  for (int t = 0; t < 200000; ++t) {
    std::vector<char> d = valid[t & 1];
    switch (t % 3) {
      case 0: d.resize(rng() % (d.size() + 1)); break;
      case 1: for (int k = 0; k < 4; ++k) d[rng() % d.size()] ^= char(1 << (rng() % 8)); break;
      default: d.resize(rng() % 300); for (char& x : d) x = char(rng()); if (d.size() >= 4) { int ty = 1 + rng() % 8; std::memcpy(d.data(), &ty, 4); } break;
    }
    // exact-size heap copy so that ASan sees any read past the end
    char* heap = new char[d.size() ? d.size() : 1];
    std::memcpy(heap, d.data(), d.size());
    RobotMessage* m = f.fromCharArray(heap, d.size());
    if (m) { ++ok; delete m; } else ++bad;
    delete[] heap;
  }
  printf("accepted %ld refused %ld\n", ok, bad);
  return 0;
}
