// TEST: the reference's C++ demo (CABAC/SimpleCABAC.cpp:48-177: 6 context bins on ctx0 init(0,20),
// 5 bypass bins, encodeBinsEP(18,5), 6 context bins on the default ctx1; then the mirror decode)
// written against include/SimpleCABAC.hpp.  Prints the stream as hex; exit code 0 when every
// decoded value matches.
#include <cstdio>

#include "SimpleCABAC.hpp"

int main(int argc, char** argv) {
  try {
    SimpleCABAC c(argc > 1 ? argv[1] : nullptr);
    const double init[6] = {0, 0, 20, 1, 1, 0};   // [ctxIdx mps state] x 2
    c.initByState(init, 2);
    const unsigned a[6] = {0, 0, 1, 0, 1, 1}, ep[5] = {1, 0, 0, 1, 0}, b[6] = {1, 1, 0, 1, 1, 1};
    c.start();
    for (unsigned v : a) c.encodeBin(v, 0);
    for (unsigned v : ep) c.encodeBinEP(v);
    c.encodeBinsEP(18, 5);
    for (unsigned v : b) c.encodeBin(v, 1);
    c.finish();
    for (unsigned char x : c.bytes()) printf("%02x", x);
    printf("\n");
    int bad = 0;
    unsigned v = 0;
    c.decodeStart();
    for (unsigned w : a) { c.decodeBin(v, 0); bad += v != w; }
    for (unsigned w : ep) { c.decodeBinEP(v); bad += v != w; }
    c.decodeBinsEP(v, 5); bad += v != 18;
    for (unsigned w : b) { c.decodeBin(v, 1); bad += v != w; }
    c.decodeFinish();
    return bad;
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 100;
  }
}
