/*
 * luf_oracle.h -- CPU restatement of the reference's Monte-Carlo decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle and the CPU baseline of
 * bench.py; nothing in the product package (localuf_b200/) links, imports or
 * calls it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.
 *
 * Parity status: PINNED.  Every function is checked (tests/test_oracle_*.py)
 * against golden vectors exported from the real reference by
 * oracle/gen_golden.py (the .npz files under tests/golden).
 *
 * Conventions (same as localuf_b200/_graph.py): node index = rank under
 * (boundary first, reference index_to_id); edge index = position in
 * code.EDGES; bit k of little-endian 32-bit word w is element 32 w + k;
 * detector bit k is node B + k.
 */
#ifndef LUF_ORACLE_H
#define LUF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t n_nodes, n_boundary, n_edges, d, height;
    const int32_t *edge_u, *edge_v;                 /* [E] */
    const uint8_t *node_flags;                      /* [N] bit0 boundary, bit1 west, bit2 east, bit3 future */
    const int32_t *node_id;                         /* [N] reference index_to_id */
    const int32_t *node_long;                       /* [N] coordinate along the long axis (-1 .. d-1) */
    const int32_t *inc_off, *inc_edge, *inc_other;  /* CSR, ascending edge id per node */
    const uint32_t *west_edge_mask;                 /* [W(E)] */
    const uint8_t *edge_class;                      /* [E] 255 = never sampled */
    int32_t n_classes;
    int32_t n_dirs;                                 /* CA neighbour table width K */
    const int32_t *nbr_edge, *nbr_node;             /* [N*K] */
    const uint8_t *nbr_owned;                       /* [N*K] */
    const int32_t *actis_span;   /* [N] per-node countdown start of Actis (decoders/luf/main.py:970-976) */
    const int32_t *actis_relay;  /* [N] neighbour towards node id 0 (decoders/luf/main.py:1138-1163), -1 = controller */
    int32_t actis_global_span;   /* controller countdown start (decoders/luf/main.py:517-523) */
    int32_t noise_kind;                             /* 0 code capacity, 1 phenomenological, 2 circuit-level */
} oracle_graph;

enum { ORACLE_UF_DYNAMIC = 1, ORACLE_UF_WEST = 2, ORACLE_UF_NODE = 4 /* NodeUF, node_uf.py:45-67 */,
       ORACLE_UF_BUCKET = 8 /* BUF, buf.py:64-78 (shortest vision first); with NODE: NodeBUF, node_buf.py:35-61
                               (smallest clusters first) */ };
enum { ORACLE_DEC_UF = 0, ORACLE_DEC_MACAR = 1, ORACLE_DEC_ACTIS = 2 };

/* Bernoulli(class_p[edge_class[e]]) per edge in edge order, as
 * noise/main.py:114-115,309-314 do with random.random() (here: xoshiro256**,
 * 53-bit doubles).  rng is 4 words of state, advanced in place. */
void oracle_sample(const oracle_graph *g, const double *class_p, uint64_t rng[4], uint32_t *err_bits);

/* _base_classes.py:427-450: XOR the endpoints of every flipped edge, drop boundary nodes. */
void oracle_syndrome(const oracle_graph *g, const uint32_t *err_bits, uint32_t *syn_bits);

/* decoders/uf.py:98-344 in canonical order (oracle/ref_canonical.py).
 * growth_out[E] (optional): 0 ungrown, 1 half, 2 full, 3 burnt.  Returns the
 * number of growth rounds. */
int oracle_uf_decode(const oracle_graph *g, int flags, const uint32_t *syn_bits,
                     uint32_t *corr_bits, uint8_t *growth_out);

/* _schemes.py:123-129: parity of (error xor correction) on edges that start on the west boundary. */
int oracle_logical_batch(const oracle_graph *g, const uint32_t *err_bits, const uint32_t *corr_bits);

/* Batch decode of nshots syndromes: [nshots][W(D)] in, [nshots][W(E)] out;
 * growth_out [nshots][E] and steps [nshots][2] optional. */
int oracle_decode_batch(const oracle_graph *g, int decoder, int flags, int64_t nshots,
                        const uint32_t *syn_bits, uint32_t *corr_bits,
                        uint8_t *growth_out, int32_t *steps);

/* _schemes.py:116-155: nshots of sample -> syndrome -> decode -> logical parity on
 * nthreads host threads (independent RNG streams seeded from seed and the
 * thread index).  Returns the failure count, or -1 on error. */
int64_t oracle_mc_batch(const oracle_graph *g, int decoder, int flags, const double *class_p,
                        uint64_t seed, int64_t nshots, int nthreads);

/* ---- forward scheme (localuf/_schemes.py:282-450) on the flat window graph ---- */
typedef struct {
    int32_t c;                      /* commit height */
    const int32_t *lower_edge;      /* [E] edge lowered by c layers, -1 if it leaves the window */
    const uint32_t *commit_mask;    /* [W(E)] */
    const int32_t *art_node;        /* [E*2] detector made artificial by an endpoint on layer c, else -1 */
    const int32_t *node_lower;      /* [N] node lowered by c layers, -1 if it leaves the window */
    const uint8_t *node_side;       /* [N] 1 west, 2 east, 0 otherwise */
} oracle_forward;

/* ncycles forward-decoding cycles with the fresh errors GIVEN (bits over code edge
 * ids): init_buffer_bits = the buffer-region error before the first cycle
 * (_schemes.py:370-390).  Per cycle: decoder steps (UF: rounds, 0; Macar: tSV, tBP),
 * the commit-region leftover, and the logical errors counted. */
int oracle_forward_stream(const oracle_graph *g, const oracle_forward *f, int decoder, int flags, int32_t ncycles,
                          const uint32_t *init_buffer_bits, const uint32_t *fresh_bits,
                          int32_t *steps, uint32_t *commit_bits, int32_t *m_cycle);

/* ---- streaming: Snowflake under the frugal scheme -------------------------
 * Layered window tables (see localuf_b200/_window.py for the conventions):
 * node (q, t) -> Snowflake id; edge (block tmin, local l) -> (h-1-tmin)*EL + l. */
typedef struct {
    int32_t d, h, NLb, NLd, EL, K;
    const uint8_t *nb_valid;                        /* [NL*K] */
    const int32_t *nb_q, *nb_dt, *nb_el, *nb_et;    /* [NL*K] */
    const int32_t *edge_q, *edge_dt;                /* [EL*2] endpoints (first, second), layer relative to the block */
    const uint8_t *edge_class, *edge_fb;            /* [EL] */
    const int32_t *pos_long;                        /* [NL] */
    int32_t n_classes;
    int32_t merger_slow;                            /* Snowflake(merger='slow'), snowflake/main.py:1237-1256 */
    int32_t schedule_11;                            /* Snowflake(schedule='1:1'), snowflake/main.py:1284-1305 */
} oracle_window;

/* One Frugal.run-like stream (_schemes.py:504-564,626-644) of ncycles decoding
 * cycles with the fresh errors GIVEN: fresh_bits [ncycles][W(EL)] over the local
 * edge index of the top block (the exclude_future_boundary filter already
 * applied by the caller).  Outputs per cycle: runtime with time_only='all' and
 * 'unrooting' (snowflake/main.py:340-347), the floor-sheet correction bits
 * committed at that cycle's drop [ncycles][W(EL)], and the logical errors
 * counted after the cycle (_pairs.py:90-111). */
int oracle_sf_stream(const oracle_window *w, int32_t ncycles, const uint32_t *fresh_bits,
                     int32_t *rt_all, int32_t *rt_unrooting, uint32_t *floor_bits, int32_t *m_cycle);

/* Monte Carlo: nshots independent Frugal.run(n) streams (_schemes.py:566-649)
 * with per-edge Bernoulli sampling; returns total logical errors, adds the
 * decoding cycles simulated to *cycles. */
int64_t oracle_sf_mc(const oracle_window *w, const double *class_p, uint64_t seed, int64_t nshots,
                     int32_t n, int nthreads, int64_t *cycles);

#ifdef __cplusplus
}
#endif
#endif
