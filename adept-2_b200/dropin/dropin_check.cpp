// dropin_check.cpp -- a user program written against the plain Adept-2 API (no CUDA in sight), linked
// against the drop-in library instead of the reference's libadept.  It records with the reference's own
// expression templates, replays through Stack::jacobian_forward/_reverse/compute_adjoint/
// compute_tangent_linear (now the CUDA engine) and dumps tape + results for tests/test_gpu_dropin.py,
// which compares them bit for bit with the oracle.  Modelled on test/test_thread_safe.cpp:36-113 and
// test/test_adept.cpp:147-195.
#include <adept_arrays.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using adept::adouble;
using adept::Real;

namespace {
// read access to the recorded tape for the dump (pointer-to-member through a derived class is legal)
struct Peek : public adept::Stack {
  static void dump(const adept::Stack& s, FILE* f) {
    const int64_t ns = s.*(&Peek::n_statements_), no = s.*(&Peek::n_operations_), g = s.*(&Peek::max_gradient_);
    const std::vector<adept::uIndex>& in = s.*(&Peek::independent_index_);
    const std::vector<adept::uIndex>& de = s.*(&Peek::dependent_index_);
    const int64_t ni = in.size(), nd = de.size();
    int64_t hdr[5] = {ns, no, g, ni, nd};
    fwrite(hdr, 8, 5, f);
    fwrite(s.*(&Peek::statement_), 8, ns, f);
    fwrite(s.*(&Peek::multiplier_), 8, no, f);
    fwrite(s.*(&Peek::index_), 4, no, f);
    fwrite(&in[0], 4, ni, f);
    fwrite(&de[0], 4, nd, f);
  }
};

// "Toon" advection (test/test_thread_safe.cpp:36-48), runtime size
void toon(int nx, int nt, double c, const std::vector<adouble>& q_init, std::vector<adouble>& q) {
  std::vector<adouble> flux(nx - 1);
  for (int i = 0; i < nx; i++) q[i] = q_init[i];
  for (int j = 0; j < nt; j++) {
    for (int i = 0; i < nx - 1; i++)
      flux[i] = (exp(c * log(q[i] / q[i + 1])) - 1.0) * q[i] * q[i + 1] / (q[i] - q[i + 1]);
    for (int i = 1; i < nx - 1; i++) q[i] += flux[i - 1] - flux[i];
    q[0] = q[nx - 2];
    q[nx - 1] = q[1];
  }
}
}  // namespace

int main(int argc, char** argv) {
  const int nx = argc > 1 ? std::atoi(argv[1]) : 128;
  const int nt = argc > 2 ? std::atoi(argv[2]) : 200;
  const char* path = argc > 3 ? argv[3] : "dropin_dump.bin";
  adept::Stack s;
  std::vector<adouble> q_init(nx), q(nx);
  const double pi = 4.0 * atan(1.0);
  for (int i = 0; i < nx; i++) q_init[i] = (0.5 + 0.5 * sin((i * 2.0 * pi) / (nx - 1.5))) + 1;
  s.new_recording();
  toon(nx, nt, 0.125, q_init, q);
  s.independent(&q_init[0], nx);
  s.dependent(&q[0], nx);

  std::vector<Real> jf(nx * nx), jr(nx * nx), jauto(nx * nx);
  s.jacobian_forward(&jf[0]);
  s.jacobian_reverse(&jr[0]);
  adept::Matrix jm = s.jacobian();            // Array<2> overload: (n_dependent, n_independent)
  for (int i = 0; i < nx; i++)
    for (int j = 0; j < nx; j++) jauto[j * nx + i] = jm(i, j);

  // adjoint of sum_i w_i q_i and tangent linear along v (test/test_adept.cpp style)
  std::vector<Real> gad(nx), gtl(nx);
  for (int i = 0; i < nx; i++) q[i].set_gradient(1.0 + 0.001 * i);
  s.compute_adjoint();
  for (int i = 0; i < nx; i++) gad[i] = q_init[i].get_gradient();
  s.clear_gradients();
  for (int i = 0; i < nx; i++) q_init[i].set_gradient(std::cos(0.1 * i));
  s.compute_tangent_linear();
  for (int i = 0; i < nx; i++) gtl[i] = q[i].get_gradient();

  // error behaviour must be the reference's
  int errors_ok = 0;
  s.clear_gradients();
  try { s.compute_adjoint(); } catch (const adept::gradients_not_initialized&) { errors_ok |= 1; }
  s.clear_independents();
  try { s.jacobian(&jf[0]); } catch (const adept::dependents_or_independents_not_identified&) { errors_ok |= 2; }
  s.independent(&q_init[0], nx);
  try { adept::Matrix bad(3, 3); s.jacobian(bad); } catch (const adept::size_mismatch&) { errors_ok |= 4; }

  double rmsd = 0.0;
  for (int k = 0; k < nx * nx; k++) rmsd += (jf[k] - jr[k]) * (jf[k] - jr[k]);
  rmsd = std::sqrt(rmsd / (nx * nx));
  std::printf("nx %d nt %d  RMSD(forward, reverse) = %g  errors_ok = %d\n", nx, nt, rmsd, errors_ok);

  FILE* f = std::fopen(path, "wb");
  if (!f) return 2;
  Peek::dump(s, f);
  fwrite(&jf[0], 8, jf.size(), f);
  fwrite(&jr[0], 8, jr.size(), f);
  fwrite(&jauto[0], 8, jauto.size(), f);
  fwrite(&gad[0], 8, nx, f);
  fwrite(&gtl[0], 8, nx, f);
  int64_t tail[2] = {errors_ok, (int64_t)(rmsd <= 1.0e-5)};
  fwrite(tail, 8, 2, f);
  std::fclose(f);
  return (errors_ok == 7 && rmsd <= 1.0e-5) ? 0 : 1;
}
