// Orchestration of the sibling multi-task baselines (included by api.cu):
//   MMoEModel_original  models/sequential/mmoe.py:183-297 (_build_seq_graph), :299-337 (_attention_fcn), :26-82 (mixing, losses)
//   PLEModel            models/sequential/ple.py:24-60
//   ShareBottomModel    models/sequential/sharebottom.py:160-203
// One GPU, whole tables.  The front end (two DIN attention poolings) is kernels_sibling.cu; every dense layer with its batch
// norm runs on the grouped kernels of kernels_head.cu exactly as PAMRec's stand-alone head does; the sparse backward is the
// per-table plan / segmented reduction / TF-exact Adam of kernels_optim.cu with the satisfied-only history as a third lookup.

static SparseTable sib_table(PamrecHandle h, int which /* 0 item 1 cate 2 user_long 3 user_short */) {
  SparseTable t;
  memset(&t, 0, sizeof t);
  if (which >= 2 && h->cfg.model_kind == PAMREC_MODEL_SASREC) return t;
  const std::string p = which == 0 ? "sp.item." : (which == 1 ? "sp.cate." : "sp.user.");
  t.keys = h->wi(p + "keys"); t.idx = h->wi(p + "idx"); t.skeys = h->wi(p + "skeys"); t.sidx = h->wi(p + "sidx");
  t.uidx = h->wi(p + "uidx"); t.ukeys = h->wi(p + "ukeys"); t.slot = h->wi(p + "slot");
  int* nu = h->wi("sp.nuniq");
  double* ns = h->wd("sp_normsq");
  if (which == 0) { t.width = kI; t.n_rows = h->cfg.n_items; t.w = h->buf.item_w; t.m = h->buf.item_m; t.v = h->buf.item_v;
                    t.accum = h->wf("sp.item.accum"); t.nuniq = nu; t.normsq = ns; t.l2_has0 = h->wi("sib.has0"); }
  else if (which == 1) { t.width = kC; t.n_rows = h->cfg.n_cates; t.w = h->buf.cate_w; t.m = h->buf.cate_m; t.v = h->buf.cate_v;
                         t.accum = h->wf("sp.cate.accum"); t.nuniq = nu + 1; t.normsq = ns + 1; t.l2_has0 = h->wi("sib.has0") + 1; }
  else if (which == 2) { t.width = PAMREC_USER_DIM; t.n_rows = h->cfg.n_users; t.w = h->buf.ulong_w; t.m = h->buf.ulong_m;
                         t.v = h->buf.ulong_v; t.nuniq = nu + 2; t.normsq = ns + 2; }
  else { t.width = PAMREC_USER_DIM; t.n_rows = h->cfg.n_users; t.w = h->buf.ushort_w; t.m = h->buf.ushort_m;
         t.v = h->buf.ushort_v; t.nuniq = nu + 2; t.normsq = ns + 3; }
  return t;
}

// keys of table t: satisfied-only history ids, history ids (both in "sib.ids_*", written by the gather), target ids
static int sib_plans(PamrecHandle h, const PamrecBatch* b, cudaStream_t st) {
  const int B = b->batch;
  const int64_t N = (int64_t)B * h->cfg.max_seq_len;
  void* tmp = h->ws<char>("cub_temp");
  const size_t tmp_bytes = (size_t)h->L.ws[h->L.ws_index.at("cub_temp")].numel;
  int rc = 0;
  SparseTable ti = sib_table(h, 0), tc = sib_table(h, 1), tu = sib_table(h, 2);
  rc |= launch_sparse_plan(ti, h->wi("sib.ids_item"), b->items, 2 * N, B, 1, 0, ti.n_rows, true, tmp, tmp_bytes, st);
  rc |= launch_sparse_plan(tc, h->wi("sib.ids_cate"), b->cates, 2 * N, B, 1, 0, tc.n_rows, true, tmp, tmp_bytes, st);
  if (h->cfg.model_kind != PAMREC_MODEL_SASREC)      // SASRec has no user tables (sasrec_param_spec)
    rc |= launch_sparse_plan(tu, b->users, nullptr, B, 0, 1, 0, tu.n_rows, true, tmp, tmp_bytes, st);
  return rc;
}

static int sib_forward(PamrecHandle h, const PamrecBatch* b, int training, float* pred_out, cudaStream_t st) {
  const Layout& L = h->L;
  const int B = b->batch, T = h->cfg.max_seq_len, N = B * T, nE = L.n_expert;
  const double cntN = (double)B * T, cntB = (double)B;
  BnSet* bn = h->bn;
  auto finalize = [&](std::initializer_list<int> ids, double cnt) {
    if (!training) return;
    for (int id : ids) if (bn[id].C > 0) launch_bn_finalize(bn[id], cnt, st);
  };
  float* hb = h->wf("sib.h");
  float* tgt = h->wf("tgt");
  float* feat = h->wf("sib.feat");
  launch_sib_gather(b->satisfied_item_history, b->satisfied_cate_history, b->item_history, b->item_cate_history, b->items, b->cates,
                    h->buf.item_w, h->buf.cate_w, hb, tgt, training ? h->wi("sib.ids_item") : nullptr,
                    training ? h->wi("sib.ids_cate") : nullptr, B, T, st);
  if (training) {
    launch_sib_has0(b->item_history, b->item_cate_history, b->items, b->cates, B, T, h->wi("sib.has0"), st);
    // the three id sorts depend on the batch only: beside the forward pass
    int prc = 0;
    h->fork(st, [&](cudaStream_t s2) {
      prc = sib_plans(h, b, s2);
      cudaEventRecord(h->ev_plan, s2);
    });
    if (prc) return fail(h, "cub sort failed");
    h->plan_for = b->item_history; h->plan_rows = B;
  } else {
    for (int i = 0; i < BN_COUNT; ++i) if (bn[i].C > 0) launch_bn_eval_stat(bn[i], st);
  }
  // ---- _attention_fcn of both branches (mmoe.py:299-337): a = h . attention_mat, feature row, 80 -> 80 -> 40 -> 1 with BN over B*T rows
  for (int r = 0; r < 2; ++r) {
    DenseP a = dense_p(hb + (int64_t)r * N * kE, kE, N, 1, kE, kE, h->P(L.att_mat + r * kE * kE), 0, h->wf("sib.zero"), 0, feat + r * 80, 160);
    launch_dense_fwd(a, st);
  }
  launch_sib_feat_fwd(feat, tgt, B, T, st);
  {
    DenseP l0 = dense_p(feat, 160, N, 2, 80, 80, h->P(L.att.w0), 6400, h->P(L.att.b0), 80, h->wf("z1"), 160);
    for (int r = 0; r < 2; ++r) { l0.x_off[r] = r * 80; l0.z_off[r] = r * 80; }
    l0.out_sums = training ? bn[BN_S0].sums : nullptr;
    launch_dense_fwd(l0, st);
    finalize({BN_S0}, cntN);
    DenseP l1 = dense_p(h->wf("z1"), 160, N, 2, 80, 40, h->P(L.att.w1), 3200, h->P(L.att.b1), 40, h->wf("z2"), 80);
    for (int r = 0; r < 2; ++r) { l1.x_off[r] = r * 80; l1.z_off[r] = r * 40; }
    set_in_bn(l1, bn[BN_S0]);
    l1.out_sums = training ? bn[BN_S1].sums : nullptr;
    launch_dense_fwd(l1, st);
    finalize({BN_S1}, cntN);
    DenseP lo = dense_p(h->wf("z2"), 80, N, 2, 40, 1, h->P(L.att.wout), 40, h->P(L.att.bout), 1, h->wf("sib.score"), 2);
    for (int r = 0; r < 2; ++r) { lo.x_off[r] = r * 40; lo.z_off[r] = r; }
    set_in_bn(lo, bn[BN_S1]);
    launch_dense_fwd(lo, st);
  }
  launch_sib_pool_fwd(hb, h->wf("sib.score"), b->satisfied_mask, b->mask, tgt, h->wf("sib.aw"), h->wf("x"), B, T, st);
  // ---- mixing layer over x = long | short | target (mmoe.py:26-50, ple.py:25-59; share-bottom has none)
  const float* x = h->wf("x");
  if (nE) {
    DenseP e0 = dense_p(x, 60, B, nE, 60, 100, h->P(L.expert.w0), 6000, h->P(L.expert.b0), 100, h->wf("ze0"), nE * 100);
    for (int g = 0; g < nE; ++g) { e0.x_off[g] = 0; e0.z_off[g] = g * 100; }
    e0.out_sums = training ? bn[BN_E0].sums : nullptr;
    launch_dense_fwd(e0, st);
    DenseP g0 = dense_p(x, 60, B, 2, 60, 64, h->P(L.gate.w0), 3840, h->P(L.gate.b0), 64, h->wf("zg0"), 128);
    for (int g = 0; g < 2; ++g) { g0.x_off[g] = 0; g0.z_off[g] = g * 64; }
    g0.out_sums = training ? bn[BN_G0].sums : nullptr;
    launch_dense_fwd(g0, st);
    finalize({BN_E0, BN_G0}, cntB);
    DenseP e1 = dense_p(h->wf("ze0"), nE * 100, B, nE, 100, 64, h->P(L.expert.w1), 6400, h->P(L.expert.b1), 64, h->wf("ze1"), nE * 64);
    for (int g = 0; g < nE; ++g) { e1.x_off[g] = g * 100; e1.z_off[g] = g * 64; }
    set_in_bn(e1, bn[BN_E0]);
    e1.out_sums = training ? bn[BN_E1].sums : nullptr;
    launch_dense_fwd(e1, st);
    DenseP g1 = dense_p(h->wf("zg0"), 128, B, 2, 64, 5, h->P(L.gate.w1), 320, h->P(L.gate.b1), 5, h->wf("zg1"), 10);
    for (int g = 0; g < 2; ++g) { g1.x_off[g] = g * 64; g1.z_off[g] = g * 5; }
    set_in_bn(g1, bn[BN_G0]);
    g1.out_sums = training ? bn[BN_G1].sums : nullptr;
    launch_dense_fwd(g1, st);
    finalize({BN_E1, BN_G1}, cntB);
    launch_sib_mix_fwd(h->wf("ze1"), h->wf("zg1"), bn[BN_E1], bn[BN_G1], tgt, h->wf("u"), nE, L.gate_sel, B, st);
  }
  // ---- towers: logit_fcn on (main | target), valid_logit_fcn on (sub | target)   mmoe.py:175-179, 231-232
  {
    const int tin = L.tower_in;
    DenseP t0 = nE ? dense_p(h->wf("u"), 168, B, 2, tin, 100, h->P(L.tower.w0), tin * 100, h->P(L.tower.b0), 100, h->wf("zt0"), 200)
                   : dense_p(x, 60, B, 2, tin, 100, h->P(L.tower.w0), tin * 100, h->P(L.tower.b0), 100, h->wf("zt0"), 200);
    t0.x_off[0] = 0; t0.x_off[1] = nE ? 84 : 0;
    for (int g = 0; g < 2; ++g) t0.z_off[g] = g * 100;
    t0.out_sums = training ? bn[BN_T0].sums : nullptr;
    launch_dense_fwd(t0, st);
    finalize({BN_T0}, cntB);
    DenseP t1 = dense_p(h->wf("zt0"), 200, B, 2, 100, 64, h->P(L.tower.w1), 6400, h->P(L.tower.b1), 64, h->wf("zt1"), 128);
    for (int g = 0; g < 2; ++g) { t1.x_off[g] = g * 100; t1.z_off[g] = g * 64; }
    set_in_bn(t1, bn[BN_T0]);
    t1.out_sums = training ? bn[BN_T1].sums : nullptr;
    launch_dense_fwd(t1, st);
    finalize({BN_T1}, cntB);
    DenseP to = dense_p(h->wf("zt1"), 128, B, 2, 64, 1, h->P(L.tower.wout), 64, h->P(L.tower.bout), 1, h->wf("logits"), 2);
    for (int g = 0; g < 2; ++g) { to.x_off[g] = g * 64; to.z_off[g] = g; }
    set_in_bn(to, bn[BN_T1]);
    launch_dense_fwd(to, st);
  }
  if (pred_out) launch_sib_pred(h->wf("logits"), pred_out, B, 2, st);
  return check_cuda(h, "forward");
}

static int sib_backward(PamrecHandle h, const PamrecBatch* b, cudaStream_t st) {
  const Layout& L = h->L;
  const int B = b->batch, T = h->cfg.max_seq_len, N = B * T, nE = L.n_expert;
  const double cntN = (double)B * T, cntB = (double)B;
  BnSet* bn = h->bn;
  const float* Pb = h->buf.dense_param;
  cudaMemsetAsync(h->buf.dense_grad, 0, (size_t)L.dense_numel * 4, st);
  cudaMemsetAsync(h->wd("loss_acc"), 0, 8 * sizeof(double), st);
  cudaMemsetAsync(h->wd("sp_normsq"), 0, 8 * sizeof(double), st);
  cudaMemsetAsync(h->wd("bn.bsums"), 0, (size_t)L.ws[L.ws_index.at("bn.bsums")].numel * sizeof(double), st);
  launch_sib_loss(h->wf("logits"), b->labels_satisfied, b->labels_play, h->wf("d_logits"), h->wd("loss_acc"), B, 0.5f, 2, st);
  auto grad_of = [&](int id, const float* Z, double cnt) {
    BnGrad g; g.Z = Z; g.stat = bn[id].stat; g.gamma = bn[id].gamma; g.beta = bn[id].beta; g.bsums = bn[id].bsums; g.count = cnt;
    return g;
  };
  auto out_of = [&](int id, const float* Z) {
    BnGradOut o; o.Z = Z; o.stat = bn[id].stat; o.gamma = bn[id].gamma; o.beta = bn[id].beta; o.bsums = bn[id].bsums;
    return o;
  };
  auto dw_bn = [&](DenseDwP& w, int id, const float* Z, double cnt) {
    w.g = grad_of(id, Z, cnt); w.g_dgamma = bn[id].dgamma; w.g_dbeta = bn[id].dbeta; w.g_C = bn[id].C; w.g_scale = 1.0f;
  };
  auto side_dw = [&](const DenseDwP& w) { h->fork(st, [&](cudaStream_t s2) { launch_dense_dw(w, s2); }); };
  const float* x = h->wf("x");
  const int tin = L.tower_in;
  // ---- towers
  {
    DenseDwP w = dw_p(h->wf("zt1"), 128, B, 2, 64, 1, h->wf("d_logits"), 2, h->G(L.tower.wout), 64, h->G(L.tower.bout), 1);
    for (int g = 0; g < 2; ++g) { w.x_off[g] = g * 64; w.z_off[g] = g; }
    set_in_bn_dw(w, bn[BN_T1]);
    side_dw(w);
    DenseDxP dx = dx_p(h->wf("d_logits"), 2, B, 64, Pb, h->wf("d_t1"), 128, 0);
    for (int g = 0; g < 2; ++g) dx_add(dx, g, g * 64, g, L.tower.wout + g * 64, 1);
    dx.o = out_of(BN_T1, h->wf("zt1"));
    launch_dense_dx(dx, st);
    DenseDwP w1 = dw_p(h->wf("zt0"), 200, B, 2, 100, 64, h->wf("d_t1"), 128, h->G(L.tower.w1), 6400, h->G(L.tower.b1), 64);
    for (int g = 0; g < 2; ++g) { w1.x_off[g] = g * 100; w1.z_off[g] = g * 64; }
    set_in_bn_dw(w1, bn[BN_T0]);
    dw_bn(w1, BN_T1, h->wf("zt1"), cntB);
    side_dw(w1);
    DenseDxP x1 = dx_p(h->wf("d_t1"), 128, B, 100, Pb, h->wf("d_t0"), 200, 0);
    for (int g = 0; g < 2; ++g) dx_add(x1, g, g * 100, g * 64, L.tower.w1 + (int64_t)g * 6400, 64);
    x1.g = grad_of(BN_T1, h->wf("zt1"), cntB);
    x1.o = out_of(BN_T0, h->wf("zt0"));
    launch_dense_dx(x1, st);
    DenseDwP w0 = nE ? dw_p(h->wf("u"), 168, B, 2, tin, 100, h->wf("d_t0"), 200, h->G(L.tower.w0), tin * 100, h->G(L.tower.b0), 100)
                     : dw_p(x, 60, B, 2, tin, 100, h->wf("d_t0"), 200, h->G(L.tower.w0), tin * 100, h->G(L.tower.b0), 100);
    w0.x_off[0] = 0; w0.x_off[1] = nE ? 84 : 0;
    for (int g = 0; g < 2; ++g) w0.z_off[g] = g * 100;
    dw_bn(w0, BN_T0, h->wf("zt0"), cntB);
    side_dw(w0);
    if (nE) {
      DenseDxP x0 = dx_p(h->wf("d_t0"), 200, B, 84, Pb, h->wf("d_u"), 168, 0);
      dx_add(x0, 0, 0, 0, L.tower.w0, 100);
      dx_add(x0, 1, 84, 100, L.tower.w0 + 8400, 100);
      x0.g = grad_of(BN_T0, h->wf("zt0"), cntB);
      launch_dense_dx(x0, st);
    } else {
      DenseDxP x0 = dx_p(h->wf("d_t0"), 200, B, 60, Pb, h->wf("d_x"), 60, 0);     // both towers read x itself (sharebottom.py:200-201)
      dx_add(x0, 0, 0, 0, L.tower.w0, 100);
      dx_add(x0, 0, 0, 100, L.tower.w0 + 6000, 100);
      x0.g = grad_of(BN_T0, h->wf("zt0"), cntB);
      launch_dense_dx(x0, st);
    }
  }
  // ---- mixing layer
  if (nE) {
    launch_sib_mix_bwd(h->wf("ze1"), h->wf("zg1"), bn[BN_E1], bn[BN_G1], h->wf("d_u"), h->wf("d_e1"), h->wf("d_g1"), h->wf("d_tgt"), nE,
                       L.gate_sel, B, st);
    launch_bn_bwd_stats(bn[BN_E1], h->wf("d_e1"), h->wf("ze1"), B, st);
    launch_bn_bwd_stats(bn[BN_G1], h->wf("d_g1"), h->wf("zg1"), B, st);
    DenseDwP we = dw_p(h->wf("ze0"), nE * 100, B, nE, 100, 64, h->wf("d_e1"), nE * 64, h->G(L.expert.w1), 6400, h->G(L.expert.b1), 64);
    for (int g = 0; g < nE; ++g) { we.x_off[g] = g * 100; we.z_off[g] = g * 64; }
    set_in_bn_dw(we, bn[BN_E0]);
    dw_bn(we, BN_E1, h->wf("ze1"), cntB);
    side_dw(we);
    DenseDxP xe = dx_p(h->wf("d_e1"), nE * 64, B, 100, Pb, h->wf("d_e0"), nE * 100, 0);
    for (int g = 0; g < nE; ++g) dx_add(xe, g, g * 100, g * 64, L.expert.w1 + (int64_t)g * 6400, 64);
    xe.g = grad_of(BN_E1, h->wf("ze1"), cntB);
    xe.o = out_of(BN_E0, h->wf("ze0"));
    launch_dense_dx(xe, st);
    DenseDwP wg = dw_p(h->wf("zg0"), 128, B, 2, 64, 5, h->wf("d_g1"), 10, h->G(L.gate.w1), 320, h->G(L.gate.b1), 5);
    for (int g = 0; g < 2; ++g) { wg.x_off[g] = g * 64; wg.z_off[g] = g * 5; }
    set_in_bn_dw(wg, bn[BN_G0]);
    dw_bn(wg, BN_G1, h->wf("zg1"), cntB);
    side_dw(wg);
    DenseDxP xg = dx_p(h->wf("d_g1"), 10, B, 64, Pb, h->wf("d_g0"), 128, 0);
    for (int g = 0; g < 2; ++g) dx_add(xg, g, g * 64, g * 5, L.gate.w1 + (int64_t)g * 320, 5);
    xg.g = grad_of(BN_G1, h->wf("zg1"), cntB);
    xg.o = out_of(BN_G0, h->wf("zg0"));
    launch_dense_dx(xg, st);
    DenseDwP we0 = dw_p(x, 60, B, nE, 60, 100, h->wf("d_e0"), nE * 100, h->G(L.expert.w0), 6000, h->G(L.expert.b0), 100);
    for (int g = 0; g < nE; ++g) { we0.x_off[g] = 0; we0.z_off[g] = g * 100; }
    dw_bn(we0, BN_E0, h->wf("ze0"), cntB);
    side_dw(we0);
    DenseDwP wg0 = dw_p(x, 60, B, 2, 60, 64, h->wf("d_g0"), 128, h->G(L.gate.w0), 3840, h->G(L.gate.b0), 64);
    for (int g = 0; g < 2; ++g) { wg0.x_off[g] = 0; wg0.z_off[g] = g * 64; }
    dw_bn(wg0, BN_G0, h->wf("zg0"), cntB);
    side_dw(wg0);
    DenseDxP xe0 = dx_p(h->wf("d_e0"), nE * 100, B, 60, Pb, h->wf("d_x"), 60, 0);
    for (int g = 0; g < nE; ++g) dx_add(xe0, 0, 0, g * 100, L.expert.w0 + (int64_t)g * 6000, 100);
    xe0.g = grad_of(BN_E0, h->wf("ze0"), cntB);
    launch_dense_dx(xe0, st);
    DenseDxP xg0 = dx_p(h->wf("d_g0"), 128, B, 60, Pb, h->wf("d_x"), 60, 1);
    for (int g = 0; g < 2; ++g) dx_add(xg0, 0, 0, g * 64, L.gate.w0 + (int64_t)g * 3840, 64);
    xg0.g = grad_of(BN_G0, h->wf("zg0"), cntB);
    launch_dense_dx(xg0, st);
  }
  // ---- attention pooling of both branches
  float* hb = h->wf("sib.h");
  float* dh = h->wf("sib.dh");
  float* feat = h->wf("sib.feat");
  launch_sib_pool_bwd(hb, h->wf("sib.aw"), b->satisfied_mask, b->mask, h->wf("d_x"), h->wf("sib.d_score"), dh, B, T, st);
  {
    DenseDwP wo = dw_p(h->wf("z2"), 80, N, 2, 40, 1, h->wf("sib.d_score"), 2, h->G(L.att.wout), 40, h->G(L.att.bout), 1);
    for (int r = 0; r < 2; ++r) { wo.x_off[r] = r * 40; wo.z_off[r] = r; }
    set_in_bn_dw(wo, bn[BN_S1]);
    side_dw(wo);
    DenseDxP xo = dx_p(h->wf("sib.d_score"), 2, N, 40, Pb, h->wf("sib.d_a1"), 80, 0);
    for (int r = 0; r < 2; ++r) dx_add(xo, r, r * 40, r, L.att.wout + r * 40, 1);
    xo.o = out_of(BN_S1, h->wf("z2"));
    launch_dense_dx(xo, st);
    DenseDwP w1 = dw_p(h->wf("z1"), 160, N, 2, 80, 40, h->wf("sib.d_a1"), 80, h->G(L.att.w1), 3200, h->G(L.att.b1), 40);
    for (int r = 0; r < 2; ++r) { w1.x_off[r] = r * 80; w1.z_off[r] = r * 40; }
    set_in_bn_dw(w1, bn[BN_S0]);
    dw_bn(w1, BN_S1, h->wf("z2"), cntN);
    side_dw(w1);
    DenseDxP x1 = dx_p(h->wf("sib.d_a1"), 80, N, 80, Pb, h->wf("sib.d_a0"), 160, 0);
    for (int r = 0; r < 2; ++r) dx_add(x1, r, r * 80, r * 40, L.att.w1 + (int64_t)r * 3200, 40);
    x1.g = grad_of(BN_S1, h->wf("z2"), cntN);
    x1.o = out_of(BN_S0, h->wf("z1"));
    launch_dense_dx(x1, st);
    DenseDwP w0 = dw_p(feat, 160, N, 2, 80, 80, h->wf("sib.d_a0"), 160, h->G(L.att.w0), 6400, h->G(L.att.b0), 80);
    for (int r = 0; r < 2; ++r) { w0.x_off[r] = r * 80; w0.z_off[r] = r * 80; }
    dw_bn(w0, BN_S0, h->wf("z1"), cntN);
    side_dw(w0);
    DenseDxP x0 = dx_p(h->wf("sib.d_a0"), 160, N, 80, Pb, h->wf("sib.d_feat"), 160, 0);
    for (int r = 0; r < 2; ++r) dx_add(x0, r, r * 80, r * 80, L.att.w0 + (int64_t)r * 6400, 80);
    x0.g = grad_of(BN_S0, h->wf("z1"), cntN);
    launch_dense_dx(x0, st);
  }
  float* d_att = h->wf("sib.d_att");
  launch_sib_feat_bwd(h->wf("sib.d_feat"), feat, h->wf("tgt"), d_att, h->wf("sib.dq"), B, T, st);
  for (int r = 0; r < 2; ++r) {
    const int64_t o = (int64_t)r * N * kE;
    DenseDwP wa = dw_p(hb + o, kE, N, 1, kE, kE, d_att + o, kE, h->G(L.att_mat + r * kE * kE), 0, h->wf("sib.dummy"), 0);
    side_dw(wa);
    DenseDxP xa = dx_p(d_att + o, kE, N, kE, Pb, dh + o, kE, 1);                // dh += d_att . attention_mat^T
    dx_add(xa, 0, 0, 0, L.att_mat + r * kE * kE, kE);
    launch_dense_dx(xa, st);
  }
  launch_sib_tgt_total(nE ? h->wf("d_tgt") : nullptr, h->wf("d_x"), h->wf("sib.dq"), h->wf("d_tgt_total"), B, st);
  h->join(st);
  return check_cuda(h, "backward");
}

static int sib_apply(PamrecHandle h, const PamrecBatch* b, int64_t step, cudaStream_t st) {
  if (step < 1) return fail(h, "step must be >= 1");
  const Layout& L = h->L;
  const PamrecConfig& c = h->cfg;
  const int B = b->batch;
  const int64_t N = (int64_t)B * c.max_seq_len;
  const double b1 = c.beta1, b2 = c.beta2;
  const float lr_t = (float)((double)c.learning_rate * std::sqrt(1.0 - std::pow(b2, (double)step)) / (1.0 - std::pow(b1, (double)step)));
  double* reg = h->wd("loss_acc") + 3;
  const bool planned = h->plan_for == b->item_history && h->plan_rows == B;
  h->plan_for = nullptr;
  if (planned) cudaStreamWaitEvent(st, h->ev_plan, 0);
  else if (sib_plans(h, b, st)) return fail(h, "cub sort failed");
  const float* dh = h->wf("sib.dh");
  const float* dT = h->wf("d_tgt_total");
  for (int t = 0; t < 2; ++t) {
    SparseTable tab = sib_table(h, t);
    const int col = t == 0 ? 0 : kI;
    launch_sparse_segreduce(tab, 2 * N + B, 2 * N, dh, kE, col, dT, kE, col, tab.normsq, st);
    launch_sparse_l2norm(tab, 2 * N + B, c.embed_l2, reg, st);
    launch_sparse_adam(tab, 2 * N + B, c.sparse_adam_mode, c.embed_l2, lr_t, c.beta1, c.beta2, c.epsilon, c.max_grad_norm, c.is_clip_norm, st);
    launch_slot_reset(tab, 2 * N + B, st);
  }
  if (c.model_kind != PAMREC_MODEL_SASREC) {
    SparseTable tl = sib_table(h, 2), ts = sib_table(h, 3);
    launch_sparse_l2norm(tl, B, c.embed_l2, reg, st);
    launch_sparse_l2norm(ts, B, c.embed_l2, reg, st);
    launch_sparse_adam(tl, B, c.sparse_adam_mode, c.embed_l2, lr_t, c.beta1, c.beta2, c.epsilon, c.max_grad_norm, c.is_clip_norm, st);
    launch_sparse_adam(ts, B, c.sparse_adam_mode, c.embed_l2, lr_t, c.beta1, c.beta2, c.epsilon, c.max_grad_norm, c.is_clip_norm, st);
    launch_slot_reset(tl, B, st);
  }
  const int n_seg = (int)L.dense.size();
  launch_dense_norm(h->buf.dense_param, h->buf.dense_grad, h->wi("seg_tab"), n_seg, c.layer_l2, h->wd("seg_normsq"),
                    h->wd("sp_normsq") + 4, reg, st);
  launch_dense_adam(h->buf.dense_param, h->buf.dense_grad, h->buf.dense_m, h->buf.dense_v, h->wi("seg_id"), h->wi("seg_tab"),
                    h->wd("seg_normsq"), L.dense_numel, c.layer_l2, lr_t, nullptr, c.beta1, c.beta2, c.epsilon, c.max_grad_norm,
                    c.is_clip_norm, st);
  launch_finish_losses(h->wd("loss_acc"), h->wf("losses"), nullptr, c.embed_l2, h->wd("sp_normsq"), st);
  return check_cuda(h, "apply_gradients");
}

// ================================================================================================ SASRecModel
// models/sequential/sasrec.py:16-96 (_build_sasrec), :230-330 (multihead_attention with dense Q / K / V), :100-141 (feedforward);
// kernels_sasrec.cu.  The sparse backward is the siblings': keys = satisfied-only ids (the lookups), full-history ids (zero gradient
// rows: they only make the row an L2 row, sequential_base_model.py:640-664), target ids.
static int sas_forward(PamrecHandle h, const PamrecBatch* b, int training, float* pred_out, cudaStream_t st) {
  const Layout& L = h->L;
  const int B = b->batch, T = h->cfg.max_seq_len;
  const int64_t N = (int64_t)B * T;
  const double cntB = (double)B;
  BnSet* bn = h->bn;
  float* hb = h->wf("sib.h");
  launch_sib_gather(b->satisfied_item_history, b->satisfied_cate_history, b->item_history, b->item_cate_history, b->items, b->cates,
                    h->buf.item_w, h->buf.cate_w, hb, h->wf("tgt"), training ? h->wi("sib.ids_item") : nullptr,
                    training ? h->wi("sib.ids_cate") : nullptr, B, T, st);
  if (training) {
    launch_sib_has0(b->item_history, b->item_cate_history, b->items, b->cates, B, T, h->wi("sib.has0"), st);
    int prc = 0;
    h->fork(st, [&](cudaStream_t s2) {
      prc = sib_plans(h, b, s2);
      cudaEventRecord(h->ev_plan, s2);
    });
    if (prc) return fail(h, "cub sort failed");
    h->plan_for = b->item_history; h->plan_rows = B;
  } else {
    launch_bn_eval_stat(bn[BN_T0], st);
    launch_bn_eval_stat(bn[BN_T1], st);
  }
  launch_sas_embed(hb, h->P(L.pos), h->wf("x0"), B, T, st);
  const float* xin = h->wf("x0");
  for (int k = 0; k < 2; ++k) {
    const Layout::SasBlock& o = L.sas[k];
    const std::string p = "blk" + std::to_string(k) + ".";
    launch_sas_proj_fwd(xin, h->P(o.wqkv), h->P(o.bqkv), h->P(o.ln_a_beta), h->P(o.ln_a_gamma), h->wf(p + "xq"), h->wf(p + "qkv"), N, st);
    launch_sas_attn_fwd(h->wf(p + "qkv"), h->wf(p + "xq"), b->satisfied_mask, h->wf(p + "y"), h->wf(p + "ml"), B, T, st);
    launch_sas_ffn_fwd(h->wf(p + "y"), h->P(o.w1), h->P(o.b1), h->P(o.w2), h->P(o.b2), h->P(o.ln_b_beta), h->P(o.ln_b_gamma), h->wf(p + "f"),
                       h->wf(p + "hpre"), h->wf(p + "out"), N, st);
    xin = h->wf(p + "out");
  }
  launch_sas_final_fwd(xin, b->satisfied_mask, h->wf("tgt"), h->wf("u"), B, T, st);
  {
    DenseP t0 = dense_p(h->wf("u"), 40, B, 1, 40, 100, h->P(L.tower.w0), 0, h->P(L.tower.b0), 0, h->wf("zt0"), 100);
    t0.out_sums = training ? bn[BN_T0].sums : nullptr;
    launch_dense_fwd(t0, st);
    if (training) launch_bn_finalize(bn[BN_T0], cntB, st);
    DenseP t1 = dense_p(h->wf("zt0"), 100, B, 1, 100, 64, h->P(L.tower.w1), 0, h->P(L.tower.b1), 0, h->wf("zt1"), 64);
    set_in_bn(t1, bn[BN_T0]);
    t1.out_sums = training ? bn[BN_T1].sums : nullptr;
    launch_dense_fwd(t1, st);
    if (training) launch_bn_finalize(bn[BN_T1], cntB, st);
    DenseP to = dense_p(h->wf("zt1"), 64, B, 1, 64, 1, h->P(L.tower.wout), 0, h->P(L.tower.bout), 0, h->wf("logits"), 1);
    set_in_bn(to, bn[BN_T1]);
    launch_dense_fwd(to, st);
  }
  if (pred_out) launch_sib_pred(h->wf("logits"), pred_out, B, 1, st);
  return check_cuda(h, "forward");
}

static int sas_backward(PamrecHandle h, const PamrecBatch* b, cudaStream_t st) {
  const Layout& L = h->L;
  const int B = b->batch, T = h->cfg.max_seq_len;
  const int64_t N = (int64_t)B * T;
  const double cntB = (double)B;
  BnSet* bn = h->bn;
  const float* Pb = h->buf.dense_param;
  cudaMemsetAsync(h->buf.dense_grad, 0, (size_t)L.dense_numel * 4, st);
  cudaMemsetAsync(h->wd("loss_acc"), 0, 8 * sizeof(double), st);
  cudaMemsetAsync(h->wd("sp_normsq"), 0, 8 * sizeof(double), st);
  cudaMemsetAsync(h->wd("bn.bsums"), 0, (size_t)L.ws[L.ws_index.at("bn.bsums")].numel * sizeof(double), st);
  launch_sib_loss(h->wf("logits"), b->labels_satisfied, nullptr, h->wf("d_logits"), h->wd("loss_acc"), B, 0.f, 1, st);
  auto grad_of = [&](int id, const float* Z, double cnt) {
    BnGrad g; g.Z = Z; g.stat = bn[id].stat; g.gamma = bn[id].gamma; g.beta = bn[id].beta; g.bsums = bn[id].bsums; g.count = cnt;
    return g;
  };
  auto out_of = [&](int id, const float* Z) {
    BnGradOut o; o.Z = Z; o.stat = bn[id].stat; o.gamma = bn[id].gamma; o.beta = bn[id].beta; o.bsums = bn[id].bsums;
    return o;
  };
  auto dw_bn = [&](DenseDwP& w, int id, const float* Z, double cnt) {
    w.g = grad_of(id, Z, cnt); w.g_dgamma = bn[id].dgamma; w.g_dbeta = bn[id].dbeta; w.g_C = bn[id].C; w.g_scale = 1.0f;
  };
  auto side_dw = [&](const DenseDwP& w) { h->fork(st, [&](cudaStream_t s2) { launch_dense_dw(w, s2); }); };
  // ---- tower (sequential_base_model.py:76-79 -> _fcn_net)
  {
    DenseDwP w = dw_p(h->wf("zt1"), 64, B, 1, 64, 1, h->wf("d_logits"), 1, h->G(L.tower.wout), 0, h->G(L.tower.bout), 0);
    set_in_bn_dw(w, bn[BN_T1]);
    side_dw(w);
    DenseDxP dx = dx_p(h->wf("d_logits"), 1, B, 64, Pb, h->wf("d_t1"), 64, 0);
    dx_add(dx, 0, 0, 0, L.tower.wout, 1);
    dx.o = out_of(BN_T1, h->wf("zt1"));
    launch_dense_dx(dx, st);
    DenseDwP w1 = dw_p(h->wf("zt0"), 100, B, 1, 100, 64, h->wf("d_t1"), 64, h->G(L.tower.w1), 0, h->G(L.tower.b1), 0);
    set_in_bn_dw(w1, bn[BN_T0]);
    dw_bn(w1, BN_T1, h->wf("zt1"), cntB);
    side_dw(w1);
    DenseDxP x1 = dx_p(h->wf("d_t1"), 64, B, 100, Pb, h->wf("d_t0"), 100, 0);
    dx_add(x1, 0, 0, 0, L.tower.w1, 64);
    x1.g = grad_of(BN_T1, h->wf("zt1"), cntB);
    x1.o = out_of(BN_T0, h->wf("zt0"));
    launch_dense_dx(x1, st);
    DenseDwP w0 = dw_p(h->wf("u"), 40, B, 1, 40, 100, h->wf("d_t0"), 100, h->G(L.tower.w0), 0, h->G(L.tower.b0), 0);
    dw_bn(w0, BN_T0, h->wf("zt0"), cntB);
    side_dw(w0);
    DenseDxP x0 = dx_p(h->wf("d_t0"), 100, B, 40, Pb, h->wf("d_u"), 40, 0);
    dx_add(x0, 0, 0, 0, L.tower.w0, 100);
    x0.g = grad_of(BN_T0, h->wf("zt0"), cntB);
    launch_dense_dx(x0, st);
  }
  float* g_a = h->wf("sas.g_a");
  float* g_b = h->wf("sas.g_b");
  float* dh = h->wf("sib.dh");
  launch_sas_final_bwd(h->wf("d_u"), b->satisfied_mask, g_a, h->wf("d_tgt_total"), B, T, st);
  // ---- encoder blocks (gradient of block 1's output is in g_a)
  for (int k = 1; k >= 0; --k) {
    const Layout::SasBlock& o = L.sas[k];
    const std::string p = "blk" + std::to_string(k) + ".";
    const float* gout = k == 1 ? g_a : g_b;
    float* gin = k == 1 ? g_b : dh;                          // block 0's input gradient = rows of the satisfied lookups (+ position rows)
    launch_sas_ffn_bwd(h->wf(p + "y"), h->wf(p + "hpre"), gout, h->P(o.w1), h->P(o.w2), h->P(o.ln_b_gamma), h->wf("sas.hid"),
                       h->wf("sas.d_hpre"), h->wf("sas.d_y"), h->G(o.ln_b_gamma), h->G(o.ln_b_beta), N, st);
    // weight gradients on the main stream: sas.hid / sas.d_hpre / sas.d_qkv are reused by the next block
    launch_dense_dw(dw_p(h->wf("sas.hid"), kE, (int)N, 1, kE, kE, gout, kE, h->G(o.w2), 0, h->G(o.b2), 0), st);
    launch_dense_dw(dw_p(h->wf(p + "f"), kE, (int)N, 1, kE, kE, h->wf("sas.d_hpre"), kE, h->G(o.w1), 0, h->G(o.b1), 0), st);
    launch_sas_attn_bwd(h->wf(p + "qkv"), h->wf(p + "xq"), h->wf(p + "y"), h->wf("sas.d_y"), h->wf(p + "ml"), b->satisfied_mask,
                        h->wf("sas.d_qkv"), h->wf("sas.dq"), B, T, st);
    launch_sas_proj_bwd(h->wf(p + "xq"), h->wf("sas.d_qkv"), h->wf("sas.d_y"), h->P(o.wqkv), h->P(o.ln_a_gamma), gin, h->G(o.ln_a_gamma),
                        h->G(o.ln_a_beta), N, st);
    DenseDwP wq = dw_p(h->wf(p + "xq"), 2 * kE, (int)N, 3, kE, kE, h->wf("sas.d_qkv"), 3 * kE, h->G(o.wqkv), kE * kE, h->G(o.bqkv), kE);
    wq.x_off[0] = 0; wq.x_off[1] = kE; wq.x_off[2] = kE;
    for (int g = 0; g < 3; ++g) wq.z_off[g] = g * kE;
    launch_dense_dw(wq, st);
  }
  launch_sas_pos_bwd(dh, h->G(L.pos), h->wd("sp_normsq") + 4, B, T, st);
  cudaMemsetAsync(dh + N * kE, 0, (size_t)N * kE * sizeof(float), st);      // the full-history "lookups" carry no gradient
  h->join(st);
  return check_cuda(h, "backward");
}
