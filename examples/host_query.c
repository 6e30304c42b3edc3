/* Minimal C host of the C-ABI (include/pnp_b200.h): proves that the header is plain C and that a non-Python host links
 * against dt4image_restoration_b200/csrc/libpnp_b200.so directly.  Queries only (no CUDA call), so it runs on a box
 * without a GPU; with a B200 present, `host_query --init` also brings the library up (pnp_init).
 *
 *   gcc -std=c99 -I include examples/host_query.c -o host_query \
 *       -L dt4image_restoration_b200/csrc -lpnp_b200 -Wl,-rpath,$PWD/dt4image_restoration_b200/csrc
 */
#include <stdio.h>
#include <string.h>

#include "pnp_b200.h"

int main(int argc, char** argv) {
  const int B = 64, H = 256, W = 256;
  size_t y0p = 0, maskp = 0;
  printf("abi %d\n", pnp_abi_version());
  printf("unet params %zu packed %zu workspace(B=%d,%dx%d) %zu\n", pnp_unet_num_params(), pnp_unet_packed_bytes(), B, H, W,
         pnp_unet_workspace_bytes(B, H, W));
  printf("prox workspace %zu prepared_supported %d\n", pnp_prox_workspace_bytes(B, H, W), pnp_prox_prepared_supported(H, W));
  if (pnp_prox_prepared_bytes(B, H, W, &y0p, &maskp) == 0) printf("prepared y0 %zu mask %zu\n", y0p, maskp);
  if (argc > 1 && strcmp(argv[1], "--init") == 0) {
    const int rc = pnp_init();
    printf("pnp_init %d %s\n", rc, rc ? pnp_last_error() : "ok");
    return rc != 0;
  }
  return 0;
}
