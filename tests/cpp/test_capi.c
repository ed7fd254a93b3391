/* Plain-C use of the C-ABI (include/mcig.h): what a binding in any language does. argv[1] == "api": no GPU needed. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "mcig.h"

#define CHECK(cond) do { if (!(cond)) { printf("FAILED %s:%d: %s | %s\n", __FILE__, __LINE__, #cond, mcig_last_error()); return 1; } } while (0)

int main(int argc, char ** argv)
{
    mcig_ctx * c = mcig_create(3);
    CHECK(c != NULL);
    CHECK(mcig_version() == 100);
    const int pdf = mcig_lookup_plugin(MCIG_PLUGIN_PDF, "ThreeDimGaussianPDF");
    const int obs = mcig_lookup_plugin(MCIG_PLUGIN_OBS, "XSquared");
    CHECK(pdf >= 0 && obs >= 0 && mcig_lookup_plugin(MCIG_PLUGIN_OBS, "ThreeDimGaussianPDF") < 0);
    CHECK(mcig_set_seed(c, 5649871) == MCIG_OK);
    CHECK(mcig_set_rng_mode(c, MCIG_RNG_REPLAY) == MCIG_OK);
    CHECK(mcig_add_pdf(c, pdf, NULL, 0) == MCIG_OK);
    CHECK(mcig_add_obs(c, obs, NULL, 0, /*blocksize*/ 0, /*nskip*/ 1, /*equil*/ 0, MCIG_EST_NOOP) == MCIG_OK);
    CHECK(mcig_get_nobsdim(c) == 1);
    CHECK(mcig_set_step(c, 0, 1.0) == MCIG_OK && mcig_get_step(c, 0) == 1.0);
    /* error codes instead of exceptions */
    CHECK(mcig_add_pdf(c, obs, NULL, 0) == MCIG_ERR_INVALID_ARGUMENT);
    CHECK(mcig_add_obs(c, obs, NULL, 0, 1, 1, 1, MCIG_EST_NOOP) == MCIG_ERR_INVALID_ARGUMENT);
    CHECK(strstr(mcig_last_error(), "requires estimator with error calculation") != NULL);
    CHECK(mcig_set_move(c, MCIG_MOVE_VEC, MCIG_SRRD_UNIFORM, 2, 1, NULL) == MCIG_ERR_INVALID_ARGUMENT);
    CHECK(mcig_prebuild(c) == MCIG_OK); /* JIT without a GPU */
    if (argc > 1 && strcmp(argv[1], "api") == 0) {
        printf("capi api ok\n");
        mcig_destroy(c);
        return 0;
    }
    double avg[1], err[1], x[3];
    CHECK(mcig_integrate(c, 100000, avg, err, 0, 0) == MCIG_OK);
    /* values produced by the compiled reference for this seed */
    CHECK(fabs(avg[0] - 0.49129926481208264) < 1e-12*0.5 && err[0] == 0.0);
    CHECK(mcig_get_acceptance_rate(c) == 0.50334);
    CHECK(mcig_get_x(c, 0, x) == MCIG_OK && x[1] == 0.12527902606923691);
    printf("capi gpu ok: avg %.17g acc %.5f\n", avg[0], mcig_get_acceptance_rate(c));
    mcig_destroy(c);
    return 0;
}
