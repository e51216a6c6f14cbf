// Memory layout of one agent's state and scratch in HBM (shared by host and device code).
//
// Everything is fp32, row-major, every leading dimension a multiple of 4 floats (16 B) so tiles can be
// moved with 16-byte async copies; pad columns are kept at zero (their gradients are exactly zero, so
// Adam leaves them at zero).  Concatenated inputs ([obs | action], [obs | target]) are stored as two
// blocks, each padded to 4 floats, so the second block starts 16 B aligned.
//
//   params arena (per agent):  [actor | critic_1 | critic_2 | acm | critic_1_targ | critic_2_targ | actor_targ]
//   moment arenas m, v       : same offsets, trainable prefix only (actor, critic_1, critic_2, acm)
// A net is a list of layers; a layer is W [rows x ld] (natural), W^T [ld x pad4(rows)] (the copy the forward
// GEMM reads, see gemm_tile.cuh) and b [rows padded to 4].  The fused Adam epilogue keeps both copies in step;
// target nets are only ever read through W^T (and their fc3 / bias vectors), so Polyak maintains just those.
#pragma once
#include <stdint.h>

namespace spp {

constexpr int kHidden = 256;        // SAC / DDPG nets: rltoolkit/algorithms/{sac,ddpg}/models.py
constexpr int kMaxLayers = 5;

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }

struct LayerDesc {
    int off_w;      // float offset of W [rows x ld] inside the net (natural torch layout; read by dX)
    int off_wt;     // float offset of W^T [ld x ld_t] (read by the forward); -1 for vector layers (critic fc3, BasicAcM gains)
    int ld_t;       // pad4(rows)
    int off_b;      // float offset of b inside the net
    int rows;       // out features
    int cols;       // logical in features (of the padded layout: see split)
    int ld;         // padded in features
    int split;      // if > 0: reference columns [0, split) map to [0, split), columns >= split map to
                    //         pad4(split) + (c - split)   (two-block inputs)
};

struct NetDesc {
    int n_layers;
    int size;       // floats, multiple of 32
    LayerDesc L[kMaxLayers];
};

enum NetId {
    NET_ACTOR = 0, NET_CRITIC_1 = 1, NET_CRITIC_2 = 2, NET_ACM = 3,
    NET_CRITIC_1_TARG = 4, NET_CRITIC_2_TARG = 5, NET_ACTOR_TARG = 6, NET_COUNT = 7
};

// Scratch buffers of one agent (float offsets into the agent's scratch arena).
struct ScratchDesc {
    int xo, xn;             // [B x ldo]   obs, next_obs
    int xc;                 // [B x ldc]   critic input for the critic update: [obs | action]
    int xcp;                // [B x ldc]   critic input for target / policy passes: [obs' | a]
    int xm;                 // [B x ldm]   ACM input [obs | denormalised target]
    int ha1, ha2;           // [B x 256]   actor hidden
    int ml;                 // [B x ldh]   actor head outputs (SAC: [mu | log_scale_raw]; DDPG: pre-tanh)
    int zt;                 // [B x ldo]   tanh(u) (SAC) / tanh(fc3) (DDPG), for the backward
    int epsb;               // [B x ldo]   eps used by the policy pass
    int hc1[2], hc2[2];     // [B x 256]   critic hidden (per critic)
    int dz2[2], dz1[2];     // [B x 256]   critic backward
    int hm1, hm2, hms;      // ACM hidden 1 / hidden 2 / BasicAcM skip pre-activation (fc21 x)
    int tm3;                // [B x lda]   ACM tanh(fc3) before the limit scale
    int ya, pa;             // [B x lda]   ACM training: regression target (actions_acm), prediction / d t1 terms
    int dm3, dm2, dm1, dms; // ACM backward
    int dxc;                // [B x ldc]   d loss / d critic input (policy pass)
    int dxm;                // [B x ldm]   d loss / d ACM input
    int dml;                // [B x ldh]   d loss / d actor head outputs
    int dza2, dza1;         // [B x 256]   actor backward
    int mk_hc2[2];          // relu mask of hc2 for the passes that never store hc2 (target and policy pass)
    int qpart[2];           // [B x 32] partial critic-head dot products, see EpiBiasAct::dot_out
    int mk_hc1[2], mk_ha1, mk_ha2;   // relu masks of hc1 / ha1 / ha2 as bits (64 per thread and 128 x 128 block), see EpiBiasAct::mask_out
    int vec;                // per-row vectors: r, notdone, y, logp_n, logp, q[2], dq[2]  (9 x Bp)
    int gvec;               // small gradient vectors (bias grads, fc3 grads): see kGvec*
    int size;
};

struct Layout {
    int algo, acm_kind, acm_critic;
    int ob, ac, B;
    int ldo, lda;           // pad4(ob), pad4(ac)
    int act_dim, ldact;     // critic action-block width: ac (acm_critic) or ob
    int ldc;                // ldo + ldact
    int ldm;                // 2 * ldo
    int heads, ldh;         // SAC: 2*ob ; DDPG: ob
    int hm1, hm2;           // ACM hidden sizes (64/32 or 100/50), ldm1, ldm2 padded
    int ldm1, ldm2;
    NetDesc actor, critic, acm;
    int net_off[NET_COUNT]; // float offsets into the params arena
    int params_size;        // floats per agent
    int train_size;         // floats per agent of the trainable prefix (m and v arenas)
    ScratchDesc s;
    int Bp;                 // pad4(B)
};

enum Algo { ALGO_SAC = 0, ALGO_DDPG = 1 };
enum AcmKind { ACM_MLP = 0, ACM_BASIC = 1 };

inline int add_layer(NetDesc& n, int& off, int rows, int cols, int split, bool with_transpose = true) {
    LayerDesc& l = n.L[n.n_layers];
    l.rows = rows;
    l.split = split;
    l.cols = split > 0 ? pad4(split) + (cols - split) : cols;
    l.ld = pad4(l.cols);
    l.off_w = off;
    off += rows * l.ld;
    off = (off + 31) & ~31;
    if (with_transpose) {
        l.ld_t = pad4(rows);
        l.off_wt = off;
        off += l.ld * l.ld_t;
        off = (off + 31) & ~31;
    } else {
        l.ld_t = 0;
        l.off_wt = -1;
    }
    l.off_b = off;
    off += pad4(rows);
    off = (off + 31) & ~31;
    return n.n_layers++;
}

inline Layout make_layout(int algo, int ob, int ac, int acm_kind, int acm_critic, int B) {
    Layout L{};
    L.algo = algo; L.acm_kind = acm_kind; L.acm_critic = acm_critic;
    L.ob = ob; L.ac = ac; L.B = B; L.Bp = pad4(B);
    L.ldo = pad4(ob); L.lda = pad4(ac);
    L.act_dim = acm_critic ? ac : ob;
    L.ldact = pad4(L.act_dim);
    L.ldc = L.ldo + L.ldact;
    L.ldm = 2 * L.ldo;
    L.heads = (algo == ALGO_SAC) ? 2 * ob : ob;
    L.ldh = pad4(L.heads);
    L.hm1 = acm_kind == ACM_MLP ? 64 : 100;
    L.hm2 = acm_kind == ACM_MLP ? 32 : 50;
    L.ldm1 = pad4(L.hm1); L.ldm2 = pad4(L.hm2);

    int off = 0;
    // actor: fc1, fc2, heads (SAC: fc_prob rows [0,ob) then fc_scale rows [ob,2ob); DDPG: fc3)
    add_layer(L.actor, off, kHidden, ob, 0);
    add_layer(L.actor, off, kHidden, kHidden, 0);
    add_layer(L.actor, off, L.heads, kHidden, 0);
    L.actor.size = off;
    off = 0;
    add_layer(L.critic, off, kHidden, ob + L.act_dim, ob);
    add_layer(L.critic, off, kHidden, kHidden, 0);
    add_layer(L.critic, off, 1, kHidden, 0, false);      // fc3 is used as a vector (row-wise head stages)
    L.critic.size = off;
    off = 0;
    // acm: fc1 [hm1 x (ob|ob)], fc2 [hm2 x hm1], fc3 [ac x hm2]; BasicAcM adds fc21 [hm2 x (ob|ob)] and
    // a 1-row pseudo layer holding t (col 0) and t1 (cols 4..4+ac) -- "gains".
    add_layer(L.acm, off, L.hm1, 2 * ob, ob);
    add_layer(L.acm, off, L.hm2, L.hm1, 0);
    add_layer(L.acm, off, ac, L.hm2, 0);
    if (acm_kind == ACM_BASIC) {
        add_layer(L.acm, off, L.hm2, 2 * ob, ob);
        add_layer(L.acm, off, 1, 4 + ac, 0, false);
    }
    L.acm.size = off;

    int p = 0;
    L.net_off[NET_ACTOR] = p; p += L.actor.size;
    L.net_off[NET_CRITIC_1] = p; p += L.critic.size;
    L.net_off[NET_CRITIC_2] = p; p += L.critic.size;
    L.net_off[NET_ACM] = p; p += L.acm.size;
    L.train_size = p;
    L.net_off[NET_CRITIC_1_TARG] = p; p += L.critic.size;
    L.net_off[NET_CRITIC_2_TARG] = p; p += L.critic.size;
    L.net_off[NET_ACTOR_TARG] = p; p += L.actor.size;
    L.params_size = p;

    ScratchDesc& s = L.s;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 31) & ~31; return r; };
    const int Bp = L.Bp;
    s.xo = take(Bp * L.ldo); s.xn = take(Bp * L.ldo);
    s.xc = take(Bp * L.ldc); s.xcp = take(Bp * L.ldc);
    s.xm = take(Bp * L.ldm);
    s.ha1 = take(Bp * kHidden); s.ha2 = take(Bp * kHidden);
    s.ml = take(Bp * L.ldh); s.zt = take(Bp * L.ldo); s.epsb = take(Bp * L.ldo);
    for (int i = 0; i < 2; ++i) { s.hc1[i] = take(Bp * kHidden); s.hc2[i] = take(Bp * kHidden); }
    for (int i = 0; i < 2; ++i) { s.dz2[i] = take(Bp * kHidden); s.dz1[i] = take(Bp * kHidden); }
    s.hm1 = take(Bp * L.ldm1); s.hm2 = take(Bp * L.ldm2); s.hms = take(Bp * L.ldm2);
    s.tm3 = take(Bp * L.lda);
    s.ya = take(Bp * L.lda);
    s.pa = take(Bp * L.lda);
    s.dm3 = take(Bp * L.lda); s.dm2 = take(Bp * L.ldm2); s.dm1 = take(Bp * L.ldm1); s.dms = take(Bp * L.ldm2);
    s.dxc = take(Bp * L.ldc); s.dxm = take(Bp * L.ldm); s.dml = take(Bp * L.ldh);
    s.dza2 = take(Bp * kHidden); s.dza1 = take(Bp * kHidden);
    {   // one uint64 per thread and 128 x 128 output block: ceil(B / 128) row tiles x 2 column halves x 256 threads x 2 floats
        const int mk = ((Bp + 127) / 128) * 2 * 256 * 2;
        s.mk_hc1[0] = take(mk); s.mk_hc1[1] = take(mk); s.mk_ha1 = take(mk); s.mk_ha2 = take(mk);
        s.mk_hc2[0] = take(mk); s.mk_hc2[1] = take(mk);
        s.qpart[0] = take(Bp * 32); s.qpart[1] = take(Bp * 32);
    }
    s.vec = take(9 * Bp);
    s.gvec = take(8 * 512);
    s.size = o;
    return L;
}

// row vectors inside ScratchDesc::vec (each Bp floats)
enum { VEC_R = 0, VEC_ND = 1, VEC_Y = 2, VEC_LOGPN = 3, VEC_LOGP = 4, VEC_Q0 = 5, VEC_Q1 = 6, VEC_DQ0 = 7, VEC_DQ1 = 8 };
// gradient vectors inside ScratchDesc::gvec (each 512 floats): bias-gradient accumulators
enum { GV_CB1_0 = 0, GV_CB1_1 = 1, GV_AB2 = 2, GV_AB1 = 3, GV_MISC = 4 };

}  // namespace spp
