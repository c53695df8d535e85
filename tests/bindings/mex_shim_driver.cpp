// TEST: drives isscabac_b200/mex/SimpleCABACMex_b200.cpp (compiled against the stub mex.h of
// oracle/mexstub, which stands in for MATLAB) the way cabacWrapper.m and a batch caller would.
//   argv[1] = scratch bitstream file.  Exit code 0 = all checks passed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mex.h"

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

static mxArray* str(const char* s) { mxArray* a = new mxArray; a->is_char = true; a->s = s; a->m = 1; a->n = a->s.size(); return a; }
static mxArray* vec(const std::vector<double>& v) { mxArray* a = new mxArray; a->d = v; a->m = 1; a->n = v.size(); return a; }

static std::vector<mxArray*> call(int nlhs, std::vector<mxArray*> in) {
  std::vector<mxArray*> out(4, nullptr);
  std::vector<const mxArray*> cin(in.begin(), in.end());
  mexFunction(nlhs, out.data(), (int)cin.size(), cin.data());
  return out;
}

int main(int argc, char** argv) {
  const char* fn = argc > 1 ? argv[1] : "/dev/shm/mex_shim_driver.bin";
  int bad = 0;
  try {
    // --- per-bin protocol, as cabacWrapper.m drives it (KAT K2 of SURVEY.md 4.1: ab 44 24)
    mxArray* h = call(1, {str("initByState"), str(fn), vec({0, 0, 20, 1, 1, 0})})[0];
    call(0, {str("encodeStart"), h});
    const double a[6] = {0, 0, 1, 0, 1, 1}, b[6] = {1, 1, 0, 1, 1, 1};
    for (double v : a) call(0, {str("encodeBin"), h, vec({v}), vec({0})});
    for (double v : b) call(0, {str("encodeBin"), h, vec({v}), vec({1})});
    bad += call(1, {str("getNumBits"), h})[0]->d[0] != 0.0;
    call(0, {str("encodeFinish"), h});
    bad += call(1, {str("getNumBits"), h})[0]->d[0] != 24.0;
    FILE* f = fopen(fn, "rb");
    unsigned char buf[8] = {0};
    size_t n = f ? fread(buf, 1, 8, f) : 0;
    if (f) fclose(f);
    bad += !(n == 3 && buf[0] == 0xab && buf[1] == 0x44 && buf[2] == 0x24);
    call(0, {str("decodeStart"), h});
    for (double v : a) bad += call(1, {str("decodeBin"), h, vec({0})})[0]->d[0] != v;
    for (double v : b) bad += call(1, {str("decodeBin"), h, vec({1})})[0]->d[0] != v;
    call(0, {str("decodeFinish"), h});
    // --- trace statistics (Windows builds of the reference): typed outputs, 5 x M uint8 + 128 x 128 uint32
    {
      mxArray* t = call(1, {str("initByState"), str(fn), vec({0, 1, 0})})[0];   // one context, (mps 1, state 0)
      call(0, {str("setTrace"), t, vec({1})});
      call(0, {str("encodeStart"), t});
      for (double v : {1.0, 1.0, 0.0}) call(0, {str("encodeBin"), t, vec({v}), vec({0})});
      call(0, {str("encodeFinish"), t});
      std::vector<mxArray*> st = call(2, {str("getEncoderStats"), t, vec({0})});
      // states: (1,0) -> MPS -> (1,1) -> MPS -> (1,2) -> LPS -> next_lps(state 2) ; trace state = state + 64
      const unsigned char want[10] = {1, 64, 1, 65, 1, 1, 65, 1, 66, 1};
      bad += !(st[0]->m == 5 && st[0]->n == 3 && st[0]->raw.size() == 15 && memcmp(st[0]->raw.data(), want, 10) == 0);
      bad += !(st[0]->raw[10] == 0 && st[0]->raw[11] == 66);
      const unsigned* tr = (const unsigned*)st[1]->raw.data();
      bad += !(st[1]->raw.size() == 128 * 128 * 4 && tr[64 * 128 + 65] == 1 && tr[65 * 128 + 66] == 1);
    }
    // an error must surface as mexErrMsgTxt with the reference's text
    try { call(0, {str("bogus")}); bad += 1; } catch (const MexStubError& e) { bad += std::string(e.what()).find("Invalid Command") == std::string::npos; }
    // --- batch commands: a 50 x 4 matrix, ISS profile, one stream per column
    std::vector<double> sym(200), ctx(23, 1.0), off = {0, 50, 100, 150, 200};
    unsigned s = 12345;
    for (double& v : sym) { s = s * 1664525u + 1013904223u; v = (s >> 24) % 8 < 5 ? 0 : (s >> 20) % 8; }
    mxArray* cfg = vec({1, 1, 8, 3, 27, 50});
    std::vector<mxArray*> enc = call(2, {str("encodeSymbols"), cfg, vec(sym), vec(off), vec(ctx)});
    std::vector<mxArray*> dec = call(1, {str("decodeSymbols"), cfg, enc[0], enc[1], vec(off), vec(ctx)});
    bad += dec[0]->d != sym;
    printf("%zu payload bytes for 200 symbols\n", enc[0]->d.size());
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 100;
  }
  return bad;
}
