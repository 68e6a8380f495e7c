// fvtg_pack_weights: the reference checkpoint (FlashVTG.state_dict(), fp32, reference key names -
// the weight ABI of inference.py:471 `load_state_dict(ckpt["model"], strict=True)`) -> the packed device
// weights every fvtg_* call consumes (FvtgWeights).  Host-side C++ only: a caller without Python loads a
// checkpoint through include/flashvtg_b200.h alone.  What packing does (include/flashvtg_b200.h documents
// the result):
//   * every Linear / conv as bf16 [n_out][k_pad] K-major, K padded to 64, conv taps folded into K as
//     [n_out][tap][c_in] (model.py:60-71, blocks/blocks.py:37-45,93-101)
//   * LayerNorm over the raw feature dim folded into the first projection (model.py:99-110):
//       W' = bf16(W diag(gamma)),  b' = W beta + b,  wsum = rowsum(W') for the mean correction (csrc/inproj.cu)
//   * token_type_embeddings rows folded into the second projection's bias (model.py:151-152)
//   * saliency_proj2 transposed, coord_head's 2-row output conv padded to 16 rows, scalars to host fields
// Strict like torch: a missing key, an unexpected key or a wrong element count is an error naming the key.
#include <math.h>
#include <stdlib.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace fvtg {
namespace {

constexpr int D = 256, FF = 1024, MLP_H = 128;

inline int pad64(int n) { return (n + 63) / 64 * 64; }

// fp32 -> bf16 bits, round to nearest even (what torch's .to(torch.bfloat16) does)
inline uint16_t bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);   // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
inline float bf16_round(float f) {
  const uint32_t u = static_cast<uint32_t>(bf16_bits(f)) << 16;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

struct KeySpec { std::string name; int64_t numel; };

// the reference state_dict key set with element counts (model.py:81-135, transformer.py:311-330,387-405,
// blocks/blocks.py:23-50,93-101); max_q_l only sizes the unused txt_position_embed table, so that key's
// count is not checked
std::vector<KeySpec> expected_keys(const FvtgCfg& c) {
  std::vector<KeySpec> s;
  auto add = [&](const std::string& n, int64_t k) { s.push_back({n, k}); };
  add("dummy_rep_token", c.num_dummies * D);
  add("dummy_rep_pos", c.num_dummies * D);
  add("coef", c.num_levels);
  add("x", 1);
  auto layer = [&](const std::string& p, bool in_proj) {
    if (in_proj) {
      add(p + ".self_attn.in_proj_weight", 3 * D * D);
      add(p + ".self_attn.in_proj_bias", 3 * D);
    }
    add(p + ".self_attn.out_proj.weight", D * D);
    add(p + ".self_attn.out_proj.bias", D);
    add(p + ".linear1.weight", FF * D);
    add(p + ".linear1.bias", FF);
    add(p + ".linear2.weight", D * FF);
    add(p + ".linear2.bias", D);
    add(p + ".norm1.weight", D);
    add(p + ".norm1.bias", D);
    add(p + ".norm2.weight", D);
    add(p + ".norm2.bias", D);
    add(p + ".activation.weight", 1);
  };
  for (int i = 0; i < c.t2v_layers; ++i) layer("transformer.t2v_encoder.layers." + std::to_string(i), false);
  for (int i = 0; i < c.enc_layers; ++i) layer("transformer.encoder.layers." + std::to_string(i), true);
  for (int i = 0; i < c.dummy_layers; ++i) layer("txtproj_encoder.layers." + std::to_string(i), true);
  add("txt_position_embed.position_embeddings.weight", -1);
  add("txt_position_embed.LayerNorm.weight", D);
  add("txt_position_embed.LayerNorm.bias", D);
  for (const char* n : {"saliency_proj1", "saliency_proj2"}) {
    add(std::string(n) + ".weight", D * D);
    add(std::string(n) + ".bias", D);
  }
  const std::pair<const char*, int> projs[2] = {{"input_txt_proj", c.t_dim}, {"input_vid_proj", c.v_dim}};
  for (const auto& pr : projs) {
    const std::string n = pr.first;
    add(n + ".0.LayerNorm.weight", pr.second);
    add(n + ".0.LayerNorm.bias", pr.second);
    add(n + ".0.net.1.weight", static_cast<int64_t>(D) * pr.second);
    add(n + ".0.net.1.bias", D);
    add(n + ".1.LayerNorm.weight", D);
    add(n + ".1.LayerNorm.bias", D);
    add(n + ".1.net.1.weight", D * D);
    add(n + ".1.net.1.bias", D);
  }
  add("token_type_embeddings.weight", 2 * D);
  for (int l = 1; l < c.num_levels; ++l)
    for (int j = 0; j < l; ++j) {
      const std::string p = "pyramid.blocks." + std::to_string(l) + ".";
      add(p + std::to_string(1 + 5 * j) + ".weight", D * D * 2);
      add(p + std::to_string(1 + 5 * j) + ".bias", D);
      add(p + std::to_string(3 + 5 * j) + ".weight", D);
      add(p + std::to_string(3 + 5 * j) + ".bias", D);
    }
  add("pooling.att.weight", D);
  for (const char* head : {"conf_head", "class_head"}) {
    const std::string h = head;
    for (int cv = 0; cv < c.num_conv_layers; ++cv) {
      add(h + ".convs." + std::to_string(cv) + ".weight", static_cast<int64_t>(D) * D * c.head_k);
      add(h + ".convs." + std::to_string(cv) + ".bias", D);
    }
    for (int m = 0; m < c.num_mlp_layers; ++m) {
      const int din = m == 0 ? D : MLP_H, dout = m == c.num_mlp_layers - 1 ? 1 : MLP_H;
      add(h + ".fc.layers." + std::to_string(m) + ".weight", static_cast<int64_t>(dout) * din);
      add(h + ".fc.layers." + std::to_string(m) + ".bias", dout);
    }
  }
  add("coord_head.module.1.weight", static_cast<int64_t>(D) * D * c.coord_k);
  add("coord_head.module.1.bias", D);
  add("coord_head.module.3.weight", static_cast<int64_t>(2) * D * c.coord_k);
  add("coord_head.module.3.bias", 2);
  return s;
}

// Sequential carve of the packed buffer: the same walk measures (host == nullptr), fills the host staging
// copy and yields the device addresses.
struct Packer {
  uint8_t* host;
  uint8_t* dev;
  size_t off = 0;
  template <typename T>
  std::pair<T*, const T*> take(size_t n) {
    off = round_up_sz(off, 256);
    T* h = host ? reinterpret_cast<T*>(host + off) : nullptr;
    const T* d = reinterpret_cast<const T*>(dev + off);
    off += n * sizeof(T);
    return {h, d};
  }
};

struct Ctx {
  const FvtgCfg& c;
  std::unordered_map<std::string, const FvtgParam*> map;
  Packer pk;
  const float* get(const std::string& k) const {
    auto it = map.find(k);
    return it == map.end() ? nullptr : it->second->data;   // dry run: no params
  }
  bool dry() const { return pk.host == nullptr; }

  const float* f32(const float* src, size_t n) {
    auto p = pk.take<float>(n);
    if (p.first) memcpy(p.first, src, n * sizeof(float));
    return p.second;
  }
  const float* f32v(const std::vector<float>& v) {
    auto p = pk.take<float>(v.size());
    if (p.first) memcpy(p.first, v.data(), v.size() * sizeof(float));
    return p.second;
  }
  // bf16 [n_pad][k_pad] from a row accessor w(n, k) over [n][k]
  template <typename F>
  const void* b16(int n, int k, int n_pad, int k_pad, F w) {
    auto p = pk.take<uint16_t>(static_cast<size_t>(n_pad) * k_pad);
    if (p.first) {
      memset(p.first, 0, static_cast<size_t>(n_pad) * k_pad * 2);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < k; ++j) p.first[static_cast<size_t>(i) * k_pad + j] = bf16_bits(w(i, j));
    }
    return p.second;
  }
  void lin(FvtgLinear& dst, const std::string& wkey, const std::string& bkey, int n, int k) {
    const float* w = get(wkey);
    dst.w = b16(n, k, n, pad64(k), [&](int i, int j) { return w[static_cast<size_t>(i) * k + j]; });
    dst.b = f32(get(bkey), n);
  }
  // conv weight (out, in, tap) [possibly with a unit dim before tap] -> [out][tap][in]
  void conv(FvtgLinear& dst, const std::string& wkey, const std::string& bkey, int n, int taps, int n_pad) {
    const float* w = get(wkey);
    dst.w = b16(n, taps * D, n_pad, pad64(taps * D), [&](int i, int j) {
      const int t = j / D, ci = j % D;
      return w[(static_cast<size_t>(i) * D + ci) * taps + t];
    });
    std::vector<float> b(n_pad, 0.f);
    if (!dry()) memcpy(b.data(), get(bkey), n * sizeof(float));
    dst.b = f32v(b);
  }
  void ln(FvtgLN& dst, const std::string& prefix) {
    dst.g = f32(get(prefix + ".weight"), D);
    dst.b = f32(get(prefix + ".bias"), D);
  }
  void in_proj(FvtgInProj& dst, const std::string& name, int dim, int tt_row) {
    const float* g0 = get(name + ".0.LayerNorm.weight");
    const float* b0 = get(name + ".0.LayerNorm.bias");
    const float* w0 = get(name + ".0.net.1.weight");
    const float* bias0 = get(name + ".0.net.1.bias");
    dst.ln0.g = f32(g0, dim);
    dst.ln0.b = f32(b0, dim);
    // LN(x) W^T + b = rstd * ((x - m0) Wg^T - mean(x - m0) * rowsum(Wg)) + (W beta + b)   (csrc/inproj.cu)
    dst.fc0.w = b16(D, dim, D, pad64(dim), [&](int i, int j) { return w0[static_cast<size_t>(i) * dim + j] * g0[j]; });
    std::vector<float> cf(D, 0.f), ws(D, 0.f);
    if (!dry()) {
      for (int i = 0; i < D; ++i) {
        double acc = 0.0, sum = 0.0;
        for (int j = 0; j < dim; ++j) {
          const float wij = w0[static_cast<size_t>(i) * dim + j];
          acc += static_cast<double>(wij) * static_cast<double>(b0[j]);
          sum += static_cast<double>(bf16_round(wij * g0[j]));
        }
        cf[i] = static_cast<float>(acc) + bias0[i];
        ws[i] = static_cast<float>(sum);
      }
    }
    dst.fc0.b = f32v(cf);
    dst.fc0_wsum = f32v(ws);
    ln(dst.ln1, name + ".1.LayerNorm");
    const float* w1 = get(name + ".1.net.1.weight");
    dst.fc1.w = b16(D, D, D, D, [&](int i, int j) { return w1[i * D + j]; });
    std::vector<float> b1(D, 0.f);
    if (!dry()) {
      const float* bb = get(name + ".1.net.1.bias");
      const float* te = get("token_type_embeddings.weight");
      for (int i = 0; i < D; ++i) b1[i] = bb[i] + te[tt_row * D + i];   // model.py:151-152
    }
    dst.fc1.b = f32v(b1);
  }
  void layer(FvtgEncLayer& dst, const std::string& p, bool in_proj) {
    if (in_proj) lin(dst.in_proj, p + ".self_attn.in_proj_weight", p + ".self_attn.in_proj_bias", 3 * D, D);
    lin(dst.out_proj, p + ".self_attn.out_proj.weight", p + ".self_attn.out_proj.bias", D, D);
    ln(dst.norm1, p + ".norm1");
    lin(dst.ff1, p + ".linear1.weight", p + ".linear1.bias", FF, D);
    lin(dst.ff2, p + ".linear2.weight", p + ".linear2.bias", D, FF);
    ln(dst.norm2, p + ".norm2");
    dst.prelu = dry() ? 0.f : get(p + ".activation.weight")[0];
  }
  void head(FvtgScoreHead& dst, const std::string& h) {
    for (int cv = 0; cv < c.num_conv_layers; ++cv)
      conv(dst.conv[cv], h + ".convs." + std::to_string(cv) + ".weight", h + ".convs." + std::to_string(cv) + ".bias",
           D, c.head_k, D);
    for (int m = 0; m < c.num_mlp_layers - 1; ++m)
      lin(dst.mlp[m], h + ".fc.layers." + std::to_string(m) + ".weight",
          h + ".fc.layers." + std::to_string(m) + ".bias", MLP_H, m == 0 ? D : MLP_H);
    const std::string last = h + ".fc.layers." + std::to_string(c.num_mlp_layers - 1);
    dst.last_w = f32(get(last + ".weight"), MLP_H);
    dst.last_b = dry() ? 0.f : get(last + ".bias")[0];
  }

  void pack_all(FvtgWeights& w) {
    in_proj(w.vid, "input_vid_proj", c.v_dim, 1);
    in_proj(w.txt, "input_txt_proj", c.t_dim, 0);
    w.dummy_tok = f32(get("dummy_rep_token"), static_cast<size_t>(c.num_dummies) * D);
    w.dummy_pos = f32(get("dummy_rep_pos"), static_cast<size_t>(c.num_dummies) * D);
    for (int i = 0; i < c.dummy_layers; ++i) layer(w.dummy[i], "txtproj_encoder.layers." + std::to_string(i), true);
    for (int i = 0; i < c.t2v_layers; ++i) layer(w.t2v[i], "transformer.t2v_encoder.layers." + std::to_string(i), false);
    for (int i = 0; i < c.enc_layers; ++i) layer(w.enc[i], "transformer.encoder.layers." + std::to_string(i), true);
    w.sal_w1 = f32(get("saliency_proj1.weight"), D * D);
    w.sal_b1 = f32(get("saliency_proj1.bias"), D);
    {
      std::vector<float> t(D * D, 0.f);
      if (!dry()) {
        const float* s2 = get("saliency_proj2.weight");
        for (int o = 0; o < D; ++o)
          for (int i = 0; i < D; ++i) t[i * D + o] = s2[o * D + i];
      }
      w.sal_w2t = f32v(t);
    }
    w.sal_b2 = f32(get("saliency_proj2.bias"), D);
    for (int l = 1; l < c.num_levels; ++l)
      for (int j = 0; j < l; ++j) {
        const std::string p = "pyramid.blocks." + std::to_string(l) + ".";
        conv(w.pyr[l][j].conv, p + std::to_string(1 + 5 * j) + ".weight", p + std::to_string(1 + 5 * j) + ".bias", D, 2, D);
        ln(w.pyr[l][j].ln, p + std::to_string(3 + 5 * j));
      }
    head(w.cls, "class_head");
    head(w.conf, "conf_head");
    conv(w.coord1, "coord_head.module.1.weight", "coord_head.module.1.bias", D, c.coord_k, D);
    conv(w.coord2, "coord_head.module.3.weight", "coord_head.module.3.bias", 2, c.coord_k, 16);
    for (int i = 0; i < FVTG_MAX_LEVELS; ++i) w.coef[i] = (!dry() && i < c.num_levels) ? get("coef")[i] : 1.f;
    w.x = dry() ? 0.f : get("x")[0];
  }
};

int check_pack_cfg(const FvtgCfg* c) {
  if (!c) return fail(FVTG_EINVAL, "null cfg");
  if (c->abi_version != FVTG_ABI_VERSION) return fail(FVTG_EINVAL, "cfg.abi_version mismatch");
  if (c->v_dim < 1 || c->t_dim < 1 || c->num_dummies < 1 || c->num_levels < 1 || c->num_levels > FVTG_MAX_LEVELS ||
      c->dummy_layers < 0 || c->dummy_layers > FVTG_MAX_LAYERS || c->t2v_layers < 0 || c->t2v_layers > FVTG_MAX_LAYERS ||
      c->enc_layers < 0 || c->enc_layers > FVTG_MAX_LAYERS || c->num_conv_layers < 1 ||
      c->num_conv_layers > FVTG_MAX_CONVS || c->num_mlp_layers < 2 || c->num_mlp_layers > FVTG_MAX_MLP ||
      c->head_k < 1 || c->coord_k < 1)
    return fail(FVTG_EINVAL, "pack_weights: configuration out of range");
  return FVTG_OK;
}

}  // namespace
}  // namespace fvtg

using namespace fvtg;

extern "C" {

size_t fvtg_packed_weights_bytes(const FvtgCfg* cfg) {
  if (check_pack_cfg(cfg) != FVTG_OK) return 0;
  Ctx ctx{*cfg, {}, Packer{nullptr, nullptr}};
  FvtgWeights w;
  memset(&w, 0, sizeof(w));
  ctx.pack_all(w);
  return round_up_sz(ctx.pk.off, 256);
}

int32_t fvtg_pack_weights_host(const FvtgCfg* cfg, const FvtgParam* params, int32_t n_params, void* host_buf,
                               size_t host_bytes, const void* target_base, FvtgWeights* out) {
  FVTG_TRY(check_pack_cfg(cfg));
  if (!params || n_params < 1 || !host_buf || !out) return fail(FVTG_EINVAL, "pack_weights: null argument");
  if (!target_base) target_base = host_buf;
  if (reinterpret_cast<uintptr_t>(target_base) & 255)
    return fail(FVTG_EINVAL, "pack_weights: the packed buffer must be 256-byte aligned");
  Ctx ctx{*cfg, {}, Packer{nullptr, nullptr}};
  for (int i = 0; i < n_params; ++i) {
    if (!params[i].name || !params[i].data) return fail(FVTG_EINVAL, "pack_weights: parameter %d has no name / data", i);
    ctx.map[params[i].name] = &params[i];
  }
  // strict load (inference.py:471): every expected key, nothing else, right sizes
  const std::vector<KeySpec> exp = expected_keys(*cfg);
  for (const KeySpec& k : exp) {
    auto it = ctx.map.find(k.name);
    if (it == ctx.map.end()) return fail(FVTG_EINVAL, "Missing key in state_dict: '%s'", k.name.c_str());
    if (k.numel >= 0 && it->second->numel != k.numel)
      return fail(FVTG_EINVAL, "size mismatch for %s: checkpoint has %lld elements, model %lld", k.name.c_str(),
                  static_cast<long long>(it->second->numel), static_cast<long long>(k.numel));
  }
  if (ctx.map.size() != exp.size()) {
    std::unordered_map<std::string, int> known;
    for (const KeySpec& k : exp) known[k.name] = 1;
    for (const auto& kv : ctx.map)
      if (!known.count(kv.first)) return fail(FVTG_EINVAL, "Unexpected key in state_dict: '%s'", kv.first.c_str());
  }
  const size_t need = fvtg_packed_weights_bytes(cfg);
  if (need > host_bytes)
    return fail(FVTG_EWORKSPACE, "pack_weights: buffer too small: need %zu bytes, have %zu", need, host_bytes);
  memset(host_buf, 0, need);
  ctx.pk = Packer{static_cast<uint8_t*>(host_buf), const_cast<uint8_t*>(static_cast<const uint8_t*>(target_base))};
  memset(out, 0, sizeof(*out));
  ctx.pack_all(*out);
  return FVTG_OK;
}

int32_t fvtg_pack_weights(const FvtgCfg* cfg, const FvtgParam* params, int32_t n_params, void* device_buf,
                          size_t device_bytes, FvtgWeights* out, void* stream) {
  if (!device_buf) return fail(FVTG_EINVAL, "pack_weights: null device buffer");
  const size_t need = fvtg_packed_weights_bytes(cfg);
  if (need == 0) return FVTG_EINVAL;
  if (need > device_bytes)
    return fail(FVTG_EWORKSPACE, "pack_weights: buffer too small: need %zu bytes, have %zu", need, device_bytes);
  std::vector<uint8_t> staging(need);
  FVTG_TRY(fvtg_pack_weights_host(cfg, params, n_params, staging.data(), need, device_buf, out));
  // pageable staging: the call returns once the bytes are handed to the DMA engine, so `staging` may die here
  FVTG_CUDA_OK(cudaMemcpyAsync(device_buf, staging.data(), need, cudaMemcpyHostToDevice,
                               static_cast<cudaStream_t>(stream)));
  return FVTG_OK;
}

}  // extern "C"
