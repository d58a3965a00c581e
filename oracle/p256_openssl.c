/* oracle/p256_openssl.c -- TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The one arm of the reference's benchs/p256_ref.cpp that can be built here (Botan and
 * Crypto++ are absent): OpenSSL EC_POINT_mul(curve, P, NULL, randp, prv, ctx) on
 * NID_X9_62_prime256v1 (benchs/p256_ref.cpp:54-96), in a steady-clock loop because Google
 * Benchmark is absent.  usage: p256_openssl <threads> <seconds>  -> prints mults/s. */
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/obj_mac.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
typedef struct { double seconds; long count; } job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  EC_GROUP* curve = EC_GROUP_new_by_curve_name(NID_X9_62_prime256v1);
  BN_CTX* ctx = BN_CTX_new();
  BIGNUM *prv = BN_new(), *r = BN_new();
  const BIGNUM* order = EC_GROUP_get0_order(curve);
  EC_POINT *randp = EC_POINT_new(curve), *P = EC_POINT_new(curve);
  BN_rand_range(r, order);
  EC_POINT_mul(curve, randp, r, NULL, NULL, ctx); /* a random point */
  BN_rand_range(prv, order);
  long n = 0;
  double t0 = now();
  while (now() - t0 < j->seconds) {
    for (int i = 0; i < 64; i++) { EC_POINT_mul(curve, P, NULL, randp, prv, ctx); BN_add_word(prv, 1); }
    n += 64;
  }
  j->count = n;
  j->seconds = now() - t0;
  return NULL;
}

int main(int argc, char** argv) {
  int nt = argc > 1 ? atoi(argv[1]) : 1;
  double secs = argc > 2 ? atof(argv[2]) : 2.0;
  if (nt < 1) nt = 1;
  if (nt > 512) nt = 512;
  pthread_t th[512];
  job_t jobs[512];
  for (int t = 0; t < nt; t++) { jobs[t].seconds = secs; jobs[t].count = 0; pthread_create(&th[t], NULL, worker, &jobs[t]); }
  double rate = 0;
  for (int t = 0; t < nt; t++) { pthread_join(th[t], NULL); rate += jobs[t].count / jobs[t].seconds; }
  printf("{\"impl\": \"openssl EC_POINT_mul variable-base\", \"threads\": %d, \"mults_per_s\": %.1f}\n", nt, rate);
  return 0;
}
