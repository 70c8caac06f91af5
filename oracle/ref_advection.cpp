/* oracle/ref_advection.cpp -- TEST INFRASTRUCTURE (see ref_harness.cpp).
 *
 * Records the reference's OWN advection schemes (benchmark/advection_schemes.h:24-58,
 * included from $ADEPT_REF/benchmark, never copied) with real adept::adouble.
 * NX is a compile-time macro in that header (benchmark/nx.h), so oracle/Makefile
 * compiles this file once per NX with -DNX=<n> -DREF_ADV_FN=ref_adv_record_nx<n>, and
 * renames the (NX-independent, hence ODR-colliding) template names with
 * -Dtoon=toon_nx<n> -Dlax_wendroff=lax_wendroff_nx<n>.
 * The caller (ref_harness.cpp) has already activated the Stack.
 */
#include <adept.h>
#include <advection_schemes.h>

#include <cmath>
#include <vector>

extern "C" int REF_ADV_FN(int scheme, int nt, double c, const double* q0,
                          std::vector<adept::uIndex>* indep, std::vector<adept::uIndex>* dep) {
  adept::Stack* s = adept::active_stack();
  if (!s) return 20;
  static const double pi = 4.0*atan(1.0);
  // Same set-up as benchmark/differentiator.h:324-334
  std::vector<adept::aReal> q_init(NX);
  std::vector<adept::aReal> q(NX);
  std::vector<double> x(NX);
  for (int i = 0; i < NX; i++)   // benchmark/autodiff_benchmark.cpp:294
    x[i] = q0 ? q0[i] : (0.5+0.5*sin((i*2.0*pi)/(NX-1.5)))+1;
  adept::set_values(&q_init[0], NX, &x[0]);
  s->new_recording();
  if (scheme == 0) lax_wendroff(nt, c, &q_init[0], &q[0]);
  else             toon(nt, c, &q_init[0], &q[0]);
  for (int i = 0; i < NX; i++) q_init[i].push_gradient_indices(*indep);
  for (int i = 0; i < NX; i++) q[i].push_gradient_indices(*dep);
  return 0;
}
