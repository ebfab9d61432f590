// C ABI of the emitter (include/dawn_b200_codegen.h).
#include "../../../include/dawn_b200_codegen.h"
#include "codegen.hpp"
#include "proto_wire.hpp"

#include <cstdlib>
#include <cstring>
#include <string>

namespace {
thread_local std::string g_error;

b200::codegen::Options convert(const dawn_b200_options* o) {
  b200::codegen::Options r;
  if(!o)
    return r;
  r.MaxHaloSize = o->max_halo_points;
  r.UseParallelEP = o->use_parallel_ep != 0;
  r.MaxBlocksPerSM = o->max_blocks_per_sm;
  r.nsms = o->nsms;
  r.DomainSizeI = o->domain_size_i, r.DomainSizeJ = o->domain_size_j, r.DomainSizeK = o->domain_size_k;
  r.RunWithSync = o->run_with_sync != 0;
  r.BlockSizeI = o->block_size_i, r.BlockSizeJ = o->block_size_j;
  r.RowsPerThread = o->rows_per_thread, r.BlockSizeK = o->block_size_k;
  r.InlineTemporaries = o->inline_temporaries, r.FuseColumns = o->fuse_columns;
  r.UnrollK = o->unroll_k;
  r.LevelsPerThread = o->levels_per_thread, r.BlockSizeIco = o->block_size_ico;
  r.UseTMA = o->use_tma, r.TmaStages = o->tma_stages;
  r.ColumnCache = o->column_cache, r.PrefetchK = o->prefetch_k;
  r.ColumnRing = o->column_ring, r.RingLevels = o->ring_levels;
  r.ColsPerThread = o->cols_per_thread;
  r.SharedTemporaries = o->shared_temporaries;
  r.CoSchedule = o->co_schedule;
  r.RingDrain = o->ring_drain;
  r.RingEarly = o->ring_early;
  r.RingStagger = o->ring_stagger;
  r.IcoFullBlocks = o->ico_full_blocks;
  r.PairStores = o->pair_stores;
  r.IcoCacheHints = o->ico_cache_hints;
  r.IcoChain = o->ico_chain;
  r.IcoChainL1 = o->ico_chain_l1;
  r.IcoChainLag = o->ico_chain_lag;
  r.RingProducer = o->ring_producer;
  r.SinglePrecision = o->single_precision;
  r.IcoStage = o->ico_stage;
  r.IcoHoist = o->ico_hoist;
  r.RingPersistent = o->persistent_ring;
  return r;
}
} // namespace

namespace b200 {
namespace codegen {
int icoChainSize(const std::vector<iir::LocationType>& chain, bool includeCenter);
}
} // namespace b200

extern "C" {

void dawn_b200_default_options(dawn_b200_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->max_halo_points = 3;
  o->run_with_sync = 1;
  o->inline_temporaries = 1;
  o->fuse_columns = 1;
  o->use_tma = 1;
  o->column_cache = 0;
  o->column_ring = 1;
  o->shared_temporaries = 0;
  o->co_schedule = 1;
  o->ring_drain = 1;
  o->ring_early = 1;
  o->ring_producer = 0;
  o->ico_chain = 0;
  o->pair_stores = 1;
  o->ico_chain_l1 = 0;
}

int dawn_b200_parse_backend(const char* s) {
  try {
    return static_cast<int>(b200::codegen::parseBackendString(s ? s : ""));
  } catch(std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

int dawn_b200_codegen_run(const char* const* names, const char* const* iir_json,
                          const size_t* iir_len, int n, int backend, const dawn_b200_options* opts,
                          char** out_code, size_t* out_len) {
  try {
    if(!names || !iir_json || !out_code || n <= 0)
      throw std::invalid_argument("dawn_b200_codegen_run: bad arguments");
    std::map<std::string, std::string> ctx;
    for(int i = 0; i < n; ++i)
      ctx[names[i]] = iir_len ? std::string(iir_json[i], iir_len[i]) : std::string(iir_json[i]);
    auto tu = b200::codegen::run(ctx, static_cast<b200::codegen::Backend>(backend), convert(opts));
    std::string code = tu.str();
    char* buf = static_cast<char*>(std::malloc(code.size() + 1));
    if(!buf)
      throw std::bad_alloc();
    std::memcpy(buf, code.c_str(), code.size() + 1);
    *out_code = buf;
    if(out_len)
      *out_len = code.size();
    return 0;
  } catch(std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

int dawn_b200_codegen_interface(const char* iir_json, size_t iir_len, int kind, char** out_text,
                                size_t* out_len) {
  try {
    auto inst = b200::iir::parseInstantiation(iir_len ? std::string(iir_json, iir_len)
                                                      : std::string(iir_json));
    if(!inst.unstructured)
      throw std::invalid_argument("interface files are generated for unstructured instantiations");
    std::string text = kind == 0 ? b200::codegen::icoCHeader(inst) : b200::codegen::icoF90Interface(inst);
    char* buf = static_cast<char*>(std::malloc(text.size() + 1));
    if(!buf)
      throw std::bad_alloc();
    std::memcpy(buf, text.c_str(), text.size() + 1);
    *out_text = buf;
    if(out_len)
      *out_len = text.size();
    return 0;
  } catch(std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

int dawn_b200_iir_convert(const char* data, size_t len, int to_format, char** out, size_t* out_len) {
  try {
    if(!data || !out)
      throw std::invalid_argument("dawn_b200_iir_convert: bad arguments");
    std::string in(data, len);
    b200::json::Value dom = b200::proto::looksLikeJson(in)
                                ? b200::json::parse(in)
                                : b200::proto::decode(in, "StencilInstantiation");
    std::string res;
    if(to_format == DAWN_B200_IIR_JSON)
      res = b200::json::write(dom);
    else if(to_format == DAWN_B200_IIR_BYTE)
      res = b200::proto::encode(dom, "StencilInstantiation");
    else
      throw std::invalid_argument("dawn_b200_iir_convert: unknown format");
    char* buf = static_cast<char*>(std::malloc(res.size() + 1));
    if(!buf)
      throw std::bad_alloc();
    std::memcpy(buf, res.data(), res.size());
    buf[res.size()] = 0;
    *out = buf;
    if(out_len)
      *out_len = res.size();
    return 0;
  } catch(std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

int dawn_b200_iir_schema(char** out_text, size_t* out_len) {
  try {
    std::string res = b200::proto::schemaDump();
    char* buf = static_cast<char*>(std::malloc(res.size() + 1));
    if(!buf)
      throw std::bad_alloc();
    std::memcpy(buf, res.c_str(), res.size() + 1);
    *out_text = buf;
    if(out_len)
      *out_len = res.size();
    return 0;
  } catch(std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

void dawn_b200_free(void* p) { std::free(p); }
const char* dawn_b200_codegen_last_error(void) { return g_error.c_str(); }

int dawn_b200_ico_chain_size(const int* chain, int chain_len, int include_center) {
  try {
    std::vector<b200::iir::LocationType> ch;
    for(int i = 0; i < chain_len; ++i) {
      if(chain[i] < 0 || chain[i] > 2)
        return -1;
      ch.push_back(static_cast<b200::iir::LocationType>(chain[i]));
    }
    return b200::codegen::icoChainSize(ch, include_center != 0);
  } catch(std::exception& e) {
    g_error = e.what();
    return -1;
  }
}
}
