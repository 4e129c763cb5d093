// libvdb_b200_storage.so: the reference's epoch directory (format/storage.cpp) written from / loaded into the
// HBM-resident index through the public C ABI.  Arrow IPC files are produced with Apache Arrow C++ (the library the
// reference's storage layer links); reading memory-maps the file, so a list's values buffer -- already one contiguous
// row-major fp32 array -- is copied by vdb_index_append_list straight from the mapping into the list's HBM pages.
// The manifest is the reference's JSON (IndexManifest::to_json / from_json, format/storage.cpp:22-100), written and
// parsed here without jsoncpp.
#include <arrow/api.h>
#include <arrow/io/api.h>
#include <arrow/ipc/api.h>

#include <sys/stat.h>

#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "vdb_b200_storage.h"

namespace {

thread_local std::string g_err;
int32_t fail(int32_t st, const std::string& msg) {
    g_err = msg;
    return st;
}
int32_t fail_core(int32_t st, const char* what) {
    g_err = std::string(what) + ": " + vdb_last_error_string();
    return st;
}
#define CORE_TRY(expr, what)                          \
    do {                                              \
        int32_t _s = (expr);                          \
        if (_s != VDB_OK) return fail_core(_s, what); \
    } while (0)
#define ARROW_TRY(expr, what)                                                             \
    do {                                                                                  \
        arrow::Status _s = (expr);                                                        \
        if (!_s.ok()) return fail(VDB_INTERNAL, std::string(what) + ": " + _s.ToString()); \
    } while (0)

// ---------------------------------------------------------------- manifest JSON (the reference's fields)

struct ShardInfo {
    uint32_t list_id = 0;
    std::string path;
    uint64_t num_vectors = 0, file_size = 0;
};
struct Manifest {
    std::string index_name, epoch, metric;
    uint32_t dimension = 0, nlist = 0, pq_m = 0, pq_nbits = 8;
    std::vector<ShardInfo> shards;
    uint64_t created_at_ns = 0;
};

std::string json_escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        if (c == '"' || c == '\\') { o += '\\'; o += c; }
        else if (c == '\n') o += "\\n";
        else if ((unsigned char)c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o += c;
    }
    return o;
}

std::string manifest_to_json(const Manifest& m) {
    std::ostringstream o;
    o << "{\n  \"index_name\": \"" << json_escape(m.index_name) << "\",\n  \"epoch\": \"" << json_escape(m.epoch)
      << "\",\n  \"dimension\": " << m.dimension << ",\n  \"nlist\": " << m.nlist << ",\n  \"metric\": \""
      << json_escape(m.metric) << "\",\n  \"pq_params\": {\"m\": " << m.pq_m << ", \"nbits\": " << m.pq_nbits
      << "},\n  \"shards\": [";
    for (size_t i = 0; i < m.shards.size(); ++i) {
        const ShardInfo& s = m.shards[i];
        o << (i ? ",\n    " : "\n    ") << "{\"list_id\": " << s.list_id << ", \"path\": \"" << json_escape(s.path)
          << "\", \"num_vectors\": " << s.num_vectors << ", \"file_size\": " << s.file_size << "}";
    }
    o << (m.shards.empty() ? "" : "\n  ") << "],\n  \"created_at\": " << m.created_at_ns << "\n}\n";
    return o.str();
}

// a small recursive-descent JSON reader: enough for any well-formed manifest (objects, arrays, strings, numbers,
// true/false/null), values kept as text / children
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    std::string text;  // Num / Str / Bool
    std::vector<JVal> arr;
    std::map<std::string, JVal> obj;
    const JVal* get(const std::string& k) const {
        auto it = obj.find(k);
        return it == obj.end() ? nullptr : &it->second;
    }
    uint64_t u64() const { return kind == Num ? std::strtoull(text.c_str(), nullptr, 10) : 0; }
    std::string str() const { return kind == Str ? text : std::string(); }
};

struct JParser {
    const std::string& s;
    size_t i = 0;
    bool ok = true;
    explicit JParser(const std::string& src) : s(src) {}
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
    std::string string_body() {
        std::string o;
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                const char e = s[++i];
                if (e == 'n') o += '\n'; else if (e == 't') o += '\t'; else if (e == 'r') o += '\r';
                else if (e == 'b') o += '\b'; else if (e == 'f') o += '\f';
                else if (e == 'u' && i + 4 < s.size()) { o += (char)std::strtoul(s.substr(i + 1, 4).c_str(), nullptr, 16); i += 4; }
                else o += e;
                ++i;
            } else o += s[i++];
        }
        if (i >= s.size()) ok = false; else ++i;
        return o;
    }
    JVal value() {
        JVal v;
        ws();
        if (i >= s.size()) { ok = false; return v; }
        const char c = s[i];
        if (c == '{') {
            ++i; v.kind = JVal::Obj;
            if (eat('}')) return v;
            do {
                ws();
                if (!eat('"')) { ok = false; return v; }
                const std::string k = string_body();
                if (!eat(':')) { ok = false; return v; }
                v.obj[k] = value();
            } while (ok && eat(','));
            if (!eat('}')) ok = false;
        } else if (c == '[') {
            ++i; v.kind = JVal::Arr;
            if (eat(']')) return v;
            do v.arr.push_back(value()); while (ok && eat(','));
            if (!eat(']')) ok = false;
        } else if (c == '"') {
            ++i; v.kind = JVal::Str; v.text = string_body();
        } else if (!s.compare(i, 4, "true")) { v.kind = JVal::Bool; v.text = "true"; i += 4; }
        else if (!s.compare(i, 5, "false")) { v.kind = JVal::Bool; v.text = "false"; i += 5; }
        else if (!s.compare(i, 4, "null")) { i += 4; }
        else {
            const size_t b = i;
            while (i < s.size() && (std::isdigit((unsigned char)s[i]) || std::strchr("+-.eE", s[i]))) ++i;
            if (i == b) ok = false;
            v.kind = JVal::Num; v.text = s.substr(b, i - b);
        }
        return v;
    }
};

bool manifest_from_json(const std::string& text, Manifest* m) {
    JParser p(text);
    const JVal root = p.value();
    if (!p.ok || root.kind != JVal::Obj) return false;
    auto S = [&](const char* k) { const JVal* v = root.get(k); return v ? v->str() : std::string(); };
    auto U = [&](const JVal& o, const char* k) { const JVal* v = o.get(k); return v ? v->u64() : 0ull; };
    m->index_name = S("index_name"); m->epoch = S("epoch"); m->metric = S("metric");
    m->dimension = (uint32_t)U(root, "dimension"); m->nlist = (uint32_t)U(root, "nlist");
    if (const JVal* pq = root.get("pq_params")) { m->pq_m = (uint32_t)U(*pq, "m"); m->pq_nbits = (uint32_t)U(*pq, "nbits"); }
    if (const JVal* sh = root.get("shards"))
        for (const JVal& e : sh->arr) {
            ShardInfo si;
            si.list_id = (uint32_t)U(e, "list_id");
            if (const JVal* pv = e.get("path")) si.path = pv->str();
            si.num_vectors = U(e, "num_vectors"); si.file_size = U(e, "file_size");
            m->shards.push_back(si);
        }
    m->created_at_ns = U(root, "created_at");
    return true;
}

const char* metric_name(int32_t metric) {  // the strings the reference's server parses, query_service.cpp:100-108
    return metric == VDB_METRIC_IP ? "InnerProduct" : metric == VDB_METRIC_COSINE ? "Cosine" : "L2";
}

// ---------------------------------------------------------------- Arrow IPC vector files

std::shared_ptr<arrow::Schema> vector_schema() {  // create_vector_schema, format/storage.cpp:287-292
    return arrow::schema({arrow::field("id", arrow::uint64()), arrow::field("vector", arrow::list(arrow::float32()))});
}

// ArrowStorage::write_vectors (:183-226): one record batch {id, vector}; the arrays wrap the caller's memory (no
// per-element builder loop), files larger than the 2^31-element limit of list<float32> offsets are split in batches
int32_t write_vectors(const std::string& path, const float* vectors, const uint64_t* ids, uint64_t n, uint32_t dim,
                      uint64_t* file_size) {
    auto schema = vector_schema();
    auto out_r = arrow::io::FileOutputStream::Open(path);
    if (!out_r.ok()) return fail(VDB_INTERNAL, "open " + path + ": " + out_r.status().ToString());
    auto out = *out_r;
    auto wr = arrow::ipc::MakeFileWriter(out, schema);
    if (!wr.ok()) return fail(VDB_INTERNAL, "ipc writer: " + wr.status().ToString());
    auto writer = *wr;
    const uint64_t max_rows = std::max<uint64_t>(1, ((1ull << 31) - 1) / std::max<uint32_t>(dim, 1));
    for (uint64_t lo = 0; lo < n || (n == 0 && lo == 0); lo += max_rows) {
        const uint64_t m = std::min(max_rows, n - lo);
        auto id_arr = std::make_shared<arrow::UInt64Array>(
            (int64_t)m, arrow::Buffer::Wrap(ids + lo, (size_t)m));
        std::vector<int32_t> offs(m + 1);
        for (uint64_t i = 0; i <= m; ++i) offs[i] = (int32_t)(i * dim);
        auto values = std::make_shared<arrow::FloatArray>((int64_t)(m * dim),
                                                          arrow::Buffer::Wrap(vectors + lo * dim, (size_t)(m * dim)));
        auto list_arr = std::make_shared<arrow::ListArray>(arrow::list(arrow::float32()), (int64_t)m,
                                                           arrow::Buffer::Wrap(offs.data(), offs.size()), values);
        auto batch = arrow::RecordBatch::Make(schema, (int64_t)m, {id_arr, list_arr});
        ARROW_TRY(writer->WriteRecordBatch(*batch), "write batch");
        if (n == 0) break;
    }
    ARROW_TRY(writer->Close(), "close writer");
    auto pos = out->Tell();
    if (file_size) *file_size = pos.ok() ? (uint64_t)*pos : 0;
    ARROW_TRY(out->Close(), "close file");
    return VDB_OK;
}

struct MappedVectors {  // zero-copy views into the memory-mapped file, one per record batch
    std::shared_ptr<arrow::io::MemoryMappedFile> file;
    std::vector<std::shared_ptr<arrow::RecordBatch>> batches;
    struct Part { const float* v; const uint64_t* ids; uint64_t n; };
    std::vector<Part> parts;
    uint64_t n = 0;
    uint32_t dim = 0;
};

int32_t map_vectors(const std::string& path, MappedVectors* mv) {
    auto f = arrow::io::MemoryMappedFile::Open(path, arrow::io::FileMode::READ);
    if (!f.ok()) return fail(VDB_INVALID_ARGUMENT, "mmap " + path + ": " + f.status().ToString());
    mv->file = *f;
    auto rd = arrow::ipc::RecordBatchFileReader::Open(mv->file);
    if (!rd.ok()) return fail(VDB_INVALID_ARGUMENT, path + ": " + rd.status().ToString());
    auto reader = *rd;
    if (reader->schema()->num_fields() != 2 || reader->schema()->field(0)->name() != "id" ||
        reader->schema()->field(1)->name() != "vector")
        return fail(VDB_INVALID_ARGUMENT, path + ": not a vdb vector file (schema " + reader->schema()->ToString() + ")");
    for (int b = 0; b < reader->num_record_batches(); ++b) {
        auto br = reader->ReadRecordBatch(b);
        if (!br.ok()) return fail(VDB_INVALID_ARGUMENT, path + ": " + br.status().ToString());
        auto batch = *br;
        const int64_t rows = batch->num_rows();
        if (rows == 0) continue;
        auto idc = std::dynamic_pointer_cast<arrow::UInt64Array>(batch->column(0));
        auto vc = std::dynamic_pointer_cast<arrow::ListArray>(batch->column(1));
        if (!idc || !vc) return fail(VDB_INVALID_ARGUMENT, path + ": unexpected column types");
        auto vals = std::dynamic_pointer_cast<arrow::FloatArray>(vc->values());
        if (!vals) return fail(VDB_INVALID_ARGUMENT, path + ": vector values are not float32");
        const int32_t o0 = vc->value_offset(0);
        const uint32_t dim = (uint32_t)(vc->value_offset(1) - o0);
        for (int64_t i = 0; i < rows; ++i)
            if ((uint32_t)(vc->value_offset(i + 1) - vc->value_offset(i)) != dim)
                return fail(VDB_INVALID_ARGUMENT, path + ": ragged vectors");
        if (mv->dim && mv->dim != dim) return fail(VDB_INVALID_ARGUMENT, path + ": dimension changes between batches");
        mv->dim = dim;
        mv->batches.push_back(batch);
        mv->parts.push_back({vals->raw_values() + o0, idc->raw_values(), (uint64_t)rows});
        mv->n += (uint64_t)rows;
    }
    return VDB_OK;
}

std::string join(const std::string& dir, const std::string& name) {
    return (!dir.empty() && dir.back() == '/') ? dir + name : dir + "/" + name;
}

}  // namespace

extern "C" {

const char* vdb_storage_last_error(void) { return g_err.c_str(); }

int32_t vdb_storage_write_vectors(const char* path, const float* vectors, const uint64_t* ids, uint64_t n,
                                  uint32_t dim) {
    if (!path || (n && (!vectors || !ids)) || dim == 0) return fail(VDB_INVALID_ARGUMENT, "write_vectors: bad arguments");
    return write_vectors(path, vectors, ids, n, dim, nullptr);
}

int32_t vdb_storage_read_vectors(const char* path, float* vectors, uint64_t* ids, uint64_t* n, uint32_t* dim) {
    if (!path || !n || !dim) return fail(VDB_INVALID_ARGUMENT, "read_vectors: bad arguments");
    MappedVectors mv;
    const int32_t st = map_vectors(path, &mv);
    if (st != VDB_OK) return st;
    *n = mv.n;
    *dim = mv.dim;
    uint64_t at = 0;
    for (const auto& p : mv.parts) {
        if (vectors) std::memcpy(vectors + at * mv.dim, p.v, p.n * mv.dim * 4);
        if (ids) std::memcpy(ids + at, p.ids, p.n * 8);
        at += p.n;
    }
    return VDB_OK;
}

int32_t vdb_index_save_epoch(vdb_index* ix, const char* dir, const char* index_name, const char* epoch) {
    if (!ix || !dir) return fail(VDB_INVALID_ARGUMENT, "save: null argument");
    vdb_stats st;
    CORE_TRY(vdb_index_stats(ix, &st), "save");
    if (::mkdir(dir, 0777) != 0 && errno != EEXIST) return fail(VDB_INTERNAL, std::string("mkdir ") + dir + ": " + std::strerror(errno));
    Manifest m;
    m.index_name = index_name ? index_name : "";
    m.epoch = epoch ? epoch : "";
    m.dimension = st.dimension;
    m.nlist = st.nlist;
    m.metric = metric_name(st.metric);
    m.created_at_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                          std::chrono::system_clock::now().time_since_epoch()).count();
    // centroids (write_centroids: ids = centroid numbers)
    {
        std::vector<float> c((size_t)st.nlist * st.dimension);
        std::vector<uint64_t> cid(st.nlist);
        for (uint32_t i = 0; i < st.nlist; ++i) cid[i] = i;
        CORE_TRY(vdb_index_get_centroids(ix, c.data()), "save");
        const int32_t s = write_vectors(join(dir, "centroids.arrow"), c.data(), cid.data(), st.nlist, st.dimension, nullptr);
        if (s != VDB_OK) return s;
    }
    std::vector<uint64_t> sizes(st.nlist);
    CORE_TRY(vdb_index_list_sizes(ix, sizes.data()), "save");
    std::vector<float> rows;
    std::vector<uint64_t> ids;
    for (uint32_t l = 0; l < st.nlist; ++l) {
        if (!sizes[l]) continue;
        rows.resize((size_t)sizes[l] * st.dimension);
        ids.resize(sizes[l]);
        CORE_TRY(vdb_index_list_vectors(ix, l, rows.data()), "save");
        CORE_TRY(vdb_index_list_ids(ix, l, ids.data()), "save");
        ShardInfo si;
        si.list_id = l;
        si.path = "list_" + std::to_string(l) + ".arrow";
        si.num_vectors = sizes[l];
        const int32_t s = write_vectors(join(dir, si.path), rows.data(), ids.data(), sizes[l], st.dimension, &si.file_size);
        if (s != VDB_OK) return s;
        m.shards.push_back(si);
    }
    // the manifest goes last: a directory with a manifest is a complete epoch
    const std::string tmp = join(dir, "manifest.json.tmp"), fin = join(dir, "manifest.json");
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        f << manifest_to_json(m);
        if (!f) return fail(VDB_INTERNAL, "write " + tmp);
    }
    if (std::rename(tmp.c_str(), fin.c_str()) != 0) return fail(VDB_INTERNAL, "rename " + tmp);
    return VDB_OK;
}

int32_t vdb_index_save(vdb_index* ix, const char* dir) { return vdb_index_save_epoch(ix, dir, nullptr, nullptr); }

int32_t vdb_index_load(vdb_index* ix, const char* dir) {
    if (!ix || !dir) return fail(VDB_INVALID_ARGUMENT, "load: null argument");
    std::ifstream f(join(dir, "manifest.json"), std::ios::binary);
    if (!f) return fail(VDB_INVALID_ARGUMENT, std::string(dir) + ": no manifest.json (not a complete epoch)");
    std::stringstream ss;
    ss << f.rdbuf();
    Manifest m;
    if (!manifest_from_json(ss.str(), &m)) return fail(VDB_INVALID_ARGUMENT, std::string(dir) + "/manifest.json: malformed");
    vdb_stats st;
    CORE_TRY(vdb_index_stats(ix, &st), "load");
    const int32_t metric = st.metric;
    if (m.dimension != st.dimension || m.nlist != st.nlist)
        return fail(VDB_INVALID_ARGUMENT, "load: epoch is " + std::to_string(m.dimension) + "-D / nlist " +
                                              std::to_string(m.nlist) + ", the index is " + std::to_string(st.dimension) +
                                              "-D / nlist " + std::to_string(st.nlist));
    if (!m.metric.empty() && m.metric != metric_name(metric))
        return fail(VDB_INVALID_ARGUMENT, "load: epoch metric " + m.metric + " != index metric " + metric_name(metric));
    if (st.total_vectors != 0) return fail(VDB_INVALID_ARGUMENT, "load: the index is not empty");
    {
        MappedVectors c;
        const int32_t s = map_vectors(join(dir, "centroids.arrow"), &c);
        if (s != VDB_OK) return s;
        if (c.n != st.nlist || c.dim != st.dimension) return fail(VDB_INVALID_ARGUMENT, "load: centroid file has the wrong shape");
        std::vector<float> cent((size_t)st.nlist * st.dimension);
        uint64_t at = 0;
        for (const auto& p : c.parts)
            for (uint64_t i = 0; i < p.n; ++i, ++at) {
                const uint64_t id = p.ids[i];
                if (id >= st.nlist) return fail(VDB_INVALID_ARGUMENT, "load: centroid id out of range");
                std::memcpy(&cent[(size_t)id * st.dimension], p.v + i * st.dimension, (size_t)st.dimension * 4);
            }
        CORE_TRY(vdb_index_set_centroids(ix, cent.data()), "load");
    }
    std::vector<uint64_t> sizes(st.nlist, 0);
    for (const ShardInfo& si : m.shards) {
        if (si.list_id >= st.nlist) return fail(VDB_INVALID_ARGUMENT, "load: list id out of range in the manifest");
        sizes[si.list_id] += si.num_vectors;
    }
    CORE_TRY(vdb_index_balance_owners(ix, sizes.data()), "load");
    for (const ShardInfo& si : m.shards) {
        MappedVectors mv;
        const int32_t s = map_vectors(join(dir, si.path), &mv);
        if (s != VDB_OK) return s;
        if (mv.n != si.num_vectors || (mv.n && mv.dim != st.dimension))
            return fail(VDB_INVALID_ARGUMENT, "load: " + si.path + " does not match its manifest entry");
        for (const auto& p : mv.parts) CORE_TRY(vdb_index_append_list(ix, si.list_id, p.v, p.ids, p.n), "load");
    }
    CORE_TRY(vdb_index_finish_load(ix), "load");
    return VDB_OK;
}

}  // extern "C"
