// fft12_check.cpp -- CPU check of socialmediamodel_b200/csrc/fft12_core.cuh (the register-resident FFT passes of the PDE
// convolution): runs the very same pass / index functions thread by thread, phase by phase (a barrier = the end of a phase),
// compares forward and inverse transforms against a direct Float64 DFT, checks freq_of_pos / pos_of_freq, and counts
// shared-memory bank conflicts of every exchange (16-byte words, 8 lanes per wavefront).
//   g++ -O2 -std=c++17 -o /tmp/fft12_check tools/fft12_check.cpp && /tmp/fft12_check
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../socialmediamodel_b200/csrc/fft12_core.cuh"

using namespace f12;
typedef std::complex<double> cd;

static long g_conflicts = 0, g_wavefronts = 0;

// one warp-instruction: 32 consecutive tids access `words[tid]`; count extra wavefronts per quarter warp
static void account(const std::vector<int>& words) {
  for (size_t w0 = 0; w0 < words.size(); w0 += 8) {
    int cnt[8] = {0};
    for (size_t l = w0; l < w0 + 8 && l < words.size(); l++) cnt[words[l] & 7]++;
    int mx = 0;
    for (int b = 0; b < 8; b++) mx = cnt[b] > mx ? cnt[b] : mx;
    g_wavefronts += 1;
    g_conflicts += mx - 1;
  }
}

template <int P> static std::vector<float2> make_tw() {  // the pass-major table, as fast_tw_kernel builds it on the device
  std::vector<float2> tw(tw_size<P>());
  for (int e = 0; e < (int)tw.size(); e++) {
    int num, den;
    tw_exponent<P>(e, num, den);
    tw[e] = make_float2((float)cos(-2 * M_PI * num / den), (float)sin(-2 * M_PI * num / den));
  }
  return tw;
}

// closed-form shared-memory words == the generic swizzle of the positions, for every thread, slot and pass
template <int P, int PAIRS> static int check_slot_words() {
  constexpr int T = Plan<P>::T, NEX = Plan<P>::NEX;
  int bad = 0;
  for (int pass = 0; pass <= NEX; pass++)
    for (int t = 0; t < T; t++)
      for (int p = 0; p < PAIRS; p++) {
        const PassAddr ad = pass_bases<P, PAIRS>(pass, t, p);
        for (int s = 0; s < 12; s++)
          if (slot_word<P, PAIRS>(pass, s, ad) != word<PAIRS>(pass_pos<P>(pass, t, s), p)) bad++;
      }
  if (bad) printf("P=%d PAIRS=%d: %d closed-form addresses differ from the generic swizzle\n", P, PAIRS, bad);
  return bad;
}

template <int P, int PAIRS> static double run(bool inverse_dit) {
  constexpr int T = Plan<P>::T, NEX = Plan<P>::NEX, NT = T * PAIRS;
  std::vector<float2> tw = make_tw<P>();
  // input: 2 * PAIRS independent sequences
  std::vector<std::vector<cd>> x(2 * PAIRS, std::vector<cd>(P));
  srand(1 + P + PAIRS);
  for (auto& v : x) for (auto& e : v) e = cd(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
  std::vector<float4> buf0(P * PAIRS), buf1(P * PAIRS);
  float4* buf[2] = {buf0.data(), buf1.data()};
  std::vector<C2> regs(NT * 12);
  auto R = [&](int tid) -> C2(&)[12] { return *reinterpret_cast<C2(*)[12]>(&regs[tid * 12]); };
  int xch = 0;
  double err = 0;
  if (!inverse_dit) {
    // forward DIF: natural in -> permuted out
    for (int tid = 0; tid < NT; tid++) {
      const int p = tid % PAIRS, t = tid / PAIRS;
      for (int e = 0; e < 12; e++) {
        const int i = pos_r12<P>(t, e);
        R(tid)[e] = C2{make_float2((float)x[2 * p][i].real(), (float)x[2 * p + 1][i].real()), make_float2((float)x[2 * p][i].imag(), (float)x[2 * p + 1][i].imag())};
      }
    }
    for (int ph = 0; ph <= NEX; ph++) {
      for (int tid = 0; tid < NT; tid++) {
        const int p = tid % PAIRS, t = tid / PAIRS;
        if (ph > 0) load_pass<P, PAIRS>(ph, R(tid), t, p, buf[(xch - 1) & 1]);
        pass_dif<P, -1>(ph, R(tid), t, tw.data());
        if (ph < NEX) store_pass<P, PAIRS>(ph, R(tid), t, p, buf[xch & 1]);
      }
      // bank accounting of the store of pass ph and the load of pass ph+1
      if (ph < NEX)
        for (int pass : {ph, ph + 1})
          for (int s = 0; s < 12; s++)
            for (int w0 = 0; w0 < NT; w0 += 32) {
              std::vector<int> words;
              for (int tid = w0; tid < w0 + 32 && tid < NT; tid++) words.push_back(word<PAIRS>(pass_pos<P>(pass, tid / PAIRS, s), tid % PAIRS));
              account(words);
            }
      xch++;
    }
    // compare with the direct DFT at freq_of_pos
    for (int c = 0; c < 2 * PAIRS; c++) {
      std::vector<cd> X(P);
      for (int k = 0; k < P; k++) { cd s = 0; for (int n = 0; n < P; n++) s += x[c][n] * std::polar(1.0, -2 * M_PI * (double)((long)k * n % P) / P); X[k] = s; }
      for (int tid = 0; tid < NT; tid++) {
        if (tid % PAIRS != c / 2) continue;
        const int t = tid / PAIRS;
        for (int s = 0; s < 12; s++) {
          const int pos = pass_pos<P>(NEX, t, s), k = freq_of_pos<P>(pos);
          if (pos_of_freq<P>(k) != pos) { printf("pos/freq mismatch P=%d pos=%d k=%d\n", P, pos, k); exit(1); }
          const C2 v = R(tid)[s];
          const cd got = (c & 1) ? cd(v.re.y, v.im.y) : cd(v.re.x, v.im.x);
          err = fmax(err, std::abs(got - X[k]));
        }
      }
    }
  } else {
    // inverse DIT: permuted in -> natural out.  Input spectrum y in permuted order; expected x[n] = sum_k Y[k] e^{+2 pi i k n / P}
    for (int tid = 0; tid < NT; tid++) {
      const int p = tid % PAIRS, t = tid / PAIRS;
      for (int s = 0; s < 12; s++) {
        const int k = freq_of_pos<P>(pass_pos<P>(NEX, t, s));
        R(tid)[s] = C2{make_float2((float)x[2 * p][k].real(), (float)x[2 * p + 1][k].real()), make_float2((float)x[2 * p][k].imag(), (float)x[2 * p + 1][k].imag())};
      }
    }
    for (int ph = 0; ph <= NEX; ph++) {
      const int pass = NEX - ph;
      for (int tid = 0; tid < NT; tid++) {
        const int p = tid % PAIRS, t = tid / PAIRS;
        if (ph > 0) load_pass<P, PAIRS>(pass, R(tid), t, p, buf[(xch - 1) & 1]);
        pass_dit<P, +1>(pass, R(tid), t, tw.data());
        if (ph < NEX) store_pass<P, PAIRS>(pass, R(tid), t, p, buf[xch & 1]);
      }
      xch++;
    }
    for (int c = 0; c < 2 * PAIRS; c++) {
      std::vector<cd> X(P);
      for (int n = 0; n < P; n++) { cd s = 0; for (int k = 0; k < P; k++) s += x[c][k] * std::polar(1.0, 2 * M_PI * (double)((long)k * n % P) / P); X[n] = s; }
      for (int tid = 0; tid < NT; tid++) {
        if (tid % PAIRS != c / 2) continue;
        const int t = tid / PAIRS;
        for (int e = 0; e < 12; e++) {
          const C2 v = R(tid)[e];
          const cd got = (c & 1) ? cd(v.re.y, v.im.y) : cd(v.re.x, v.im.x);
          err = fmax(err, std::abs(got - X[pos_r12<P>(t, e)]));
        }
      }
    }
  }
  return err;
}

// inverse in DIF form (rows_inv): natural in -> permuted out with SGN = +1
template <int P, int PAIRS> static double run_dif_inv() {
  constexpr int T = Plan<P>::T, NEX = Plan<P>::NEX, NT = T * PAIRS;
  std::vector<float2> tw = make_tw<P>();
  std::vector<cd> x(P);
  srand(7 + P);
  for (auto& e : x) e = cd(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
  std::vector<float4> buf0(P * PAIRS), buf1(P * PAIRS);
  float4* buf[2] = {buf0.data(), buf1.data()};
  std::vector<C2> regs(NT * 12);
  auto R = [&](int tid) -> C2(&)[12] { return *reinterpret_cast<C2(*)[12]>(&regs[tid * 12]); };
  for (int tid = 0; tid < NT; tid++)
    for (int e = 0; e < 12; e++) {
      const int i = pos_r12<P>(tid / PAIRS, e);
      R(tid)[e] = C2{make_float2((float)x[i].real(), (float)x[i].real()), make_float2((float)x[i].imag(), (float)x[i].imag())};
    }
  int xch = 0;
  for (int ph = 0; ph <= NEX; ph++) {
    for (int tid = 0; tid < NT; tid++) {
      const int p = tid % PAIRS, t = tid / PAIRS;
      if (ph > 0) load_pass<P, PAIRS>(ph, R(tid), t, p, buf[(xch - 1) & 1]);
      pass_dif<P, +1>(ph, R(tid), t, tw.data());
      if (ph < NEX) store_pass<P, PAIRS>(ph, R(tid), t, p, buf[xch & 1]);
    }
    xch++;
  }
  double err = 0;
  std::vector<cd> X(P);
  for (int n = 0; n < P; n++) { cd s = 0; for (int k = 0; k < P; k++) s += x[k] * std::polar(1.0, 2 * M_PI * (double)((long)k * n % P) / P); X[n] = s; }
  for (int tid = 0; tid < NT; tid++)
    for (int s = 0; s < 12; s++) {
      const C2 v = R(tid)[s];
      err = fmax(err, std::abs(cd(v.re.x, v.im.x) - X[freq_of_pos<P>(pass_pos<P>(NEX, tid / PAIRS, s))]));
    }
  return err;
}

template <int P, int PAIRS> static int check() {
  g_conflicts = g_wavefronts = 0;
  if (check_slot_words<P, PAIRS>()) return 1;
  const double e1 = run<P, PAIRS>(false), e2 = run<P, PAIRS>(true), e3 = run_dif_inv<P, PAIRS>();
  const double scale = sqrt((double)P);  // rms magnitude of the spectrum of O(1) inputs
  printf("P=%4d PAIRS=%d: fwd err %.2e  inv(DIT) err %.2e  inv(DIF) err %.2e  (scale %.1f)  bank conflicts %ld / %ld wavefronts\n", P, PAIRS, e1, e2, e3,
         scale, g_conflicts, g_wavefronts);
  return (e1 < 2e-5 * scale && e2 < 2e-5 * scale && e3 < 2e-5 * scale && g_conflicts == 0) ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += check<96, 8>(); bad += check<192, 8>(); bad += check<384, 8>(); bad += check<768, 8>();
  bad += check<96, 4>(); bad += check<192, 4>(); bad += check<384, 4>(); bad += check<768, 4>();
  printf(bad ? "FAILED\n" : "all passes correct and conflict-free\n");
  return bad;
}
