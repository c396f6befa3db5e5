/* TEST INFRASTRUCTURE — tiny open-addressing byte-keyed hash map used by the oracle
 * (stands in for BPF_MAP_TYPE_HASH; "flat hash with the reference capacities lifted",
 * SURVEY §8(d) CPU baseline). Not thread-safe for writers. */
#ifndef CLK_ORACLE_HMAP_H
#define CLK_ORACLE_HMAP_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct hmap {
    uint32_t ksz, vsz, cap, used;   /* cap is a power of two */
    uint8_t *keys, *vals, *occ;
} hmap;

static inline uint64_t hmap_hash(const void *k, uint32_t n)
{
    const uint8_t *p = (const uint8_t *)k;
    uint64_t h = 0x9E3779B97F4A7C15ull ^ n;
    while (n >= 8) { uint64_t w; memcpy(&w, p, 8); h = (h ^ w) * 0xff51afd7ed558ccdull; h ^= h >> 32; p += 8; n -= 8; }
    if (n) { uint64_t w = 0; memcpy(&w, p, n); h = (h ^ w) * 0xff51afd7ed558ccdull; h ^= h >> 32; }
    h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 29;
    return h;
}
static inline void hmap_init(hmap *m, uint32_t ksz, uint32_t vsz, uint32_t cap_hint)
{
    uint32_t cap = 16; while (cap < cap_hint * 2) cap <<= 1;
    m->ksz = ksz; m->vsz = vsz; m->cap = cap; m->used = 0;
    m->keys = (uint8_t *)calloc(cap, ksz); m->vals = (uint8_t *)calloc(cap, vsz ? vsz : 1);
    m->occ = (uint8_t *)calloc(cap, 1);
}
static inline void hmap_free(hmap *m) { free(m->keys); free(m->vals); free(m->occ); memset(m, 0, sizeof *m); }
static inline void *hmap_get(const hmap *m, const void *k)
{
    uint32_t i = (uint32_t)hmap_hash(k, m->ksz) & (m->cap - 1);
    while (m->occ[i]) {
        if (!memcmp(m->keys + (size_t)i * m->ksz, k, m->ksz)) return m->vals + (size_t)i * m->vsz;
        i = (i + 1) & (m->cap - 1);
    }
    return NULL;
}
static inline void *hmap_put(hmap *m, const void *k, const void *v);
static inline void hmap_grow(hmap *m)
{
    hmap n; hmap_init(&n, m->ksz, m->vsz, m->cap);
    for (uint32_t i = 0; i < m->cap; i++)
        if (m->occ[i]) hmap_put(&n, m->keys + (size_t)i * m->ksz, m->vals + (size_t)i * m->vsz);
    hmap_free(m); *m = n;
}
static inline void *hmap_put(hmap *m, const void *k, const void *v)
{
    if ((m->used + 1) * 2 > m->cap) hmap_grow(m);
    uint32_t i = (uint32_t)hmap_hash(k, m->ksz) & (m->cap - 1);
    while (m->occ[i]) {
        if (!memcmp(m->keys + (size_t)i * m->ksz, k, m->ksz)) { if (v) memcpy(m->vals + (size_t)i * m->vsz, v, m->vsz); return m->vals + (size_t)i * m->vsz; }
        i = (i + 1) & (m->cap - 1);
    }
    m->occ[i] = 1; m->used++;
    memcpy(m->keys + (size_t)i * m->ksz, k, m->ksz);
    if (v) memcpy(m->vals + (size_t)i * m->vsz, v, m->vsz); else memset(m->vals + (size_t)i * m->vsz, 0, m->vsz);
    return m->vals + (size_t)i * m->vsz;
}
#endif
