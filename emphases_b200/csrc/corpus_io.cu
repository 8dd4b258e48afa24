// Native corpus ingest / egress for emphases.from_files_to_files (host code).
//
// Replaces, for a whole file list at once and on a thread pool, the per-file
// Python work of the reference loop (emphases/core.py:169-179):
//   emphases.load.audio      torchaudio.load of a wav        emphases/load.py:11-17
//   pypar.Alignment(file)    Praat TextGrid parse             emphases/core.py:49
//   alignment.save(...)      TextGrid write                   emphases/core.py:111
// 16-bit PCM samples go straight into a caller-provided (pinned) int16 buffer
// -- converted to float on the GPU, exactly x / 32768 like torchaudio -- and
// word times into a float64 array parsed with strtod (= Python float()).
// Files this reader does not understand (other encodings, UTF-16 TextGrids)
// get a non-zero status and are left to the Python path.
#include <sys/stat.h>

#include <atomic>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/emphases_b200.h"

namespace {

constexpr const char* kSilence = "<silent>";

struct Word {
    std::string label;
    double start, end;
};

struct FileEntry {
    int status = 0;                 // 0 ok
    std::string error;
    std::string audio_path;
    int sample_rate = 0, channels = 0, bits = 0, format = 0;
    long long n_samples = 0;        // per channel
    long long data_offset = 0;      // byte offset of the PCM data in the file
    std::vector<Word> words;
};

bool read_file(const std::string& path, std::string& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(size > 0 ? size : 0);
    size_t got = size > 0 ? std::fread(&out[0], 1, size, f) : 0;
    std::fclose(f);
    return got == (size_t)(size > 0 ? size : 0);
}

void fail(FileEntry& e, int status, const std::string& message) {
    if (e.status == 0) {
        e.status = status;
        e.error = message;
    }
}

// ---- wav header ----
void scan_wav(FileEntry& e) {
    FILE* f = std::fopen(e.audio_path.c_str(), "rb");
    if (!f) return fail(e, 1, "cannot open " + e.audio_path);
    unsigned char head[12];
    if (std::fread(head, 1, 12, f) != 12 || std::memcmp(head, "RIFF", 4) || std::memcmp(head + 8, "WAVE", 4)) {
        std::fclose(f);
        return fail(e, 2, "not a RIFF/WAVE file: " + e.audio_path);
    }
    bool have_fmt = false, have_data = false;
    long long cursor = 12;
    unsigned char chunk[8];
    while (std::fread(chunk, 1, 8, f) == 8) {
        unsigned size = chunk[4] | (chunk[5] << 8) | (chunk[6] << 16) | ((unsigned)chunk[7] << 24);
        cursor += 8;
        if (!std::memcmp(chunk, "fmt ", 4)) {
            unsigned char fmt[40] = {0};
            unsigned want = size < 40 ? size : 40;
            if (std::fread(fmt, 1, want, f) != want) break;
            e.format = fmt[0] | (fmt[1] << 8);
            e.channels = fmt[2] | (fmt[3] << 8);
            e.sample_rate = fmt[4] | (fmt[5] << 8) | (fmt[6] << 16) | ((unsigned)fmt[7] << 24);
            e.bits = fmt[14] | (fmt[15] << 8);
            if (e.format == 0xFFFE && size >= 26) e.format = fmt[24] | (fmt[25] << 8);
            have_fmt = true;
            std::fseek(f, cursor + size + (size & 1), SEEK_SET);
        } else if (!std::memcmp(chunk, "data", 4)) {
            e.data_offset = cursor;
            if (have_fmt && e.channels > 0 && e.bits > 0)
                e.n_samples = (long long)size / (e.channels * (e.bits / 8));
            have_data = true;
            break;
        } else {
            std::fseek(f, cursor + size + (size & 1), SEEK_SET);
        }
        cursor += size + (size & 1);
    }
    std::fclose(f);
    if (!have_fmt || !have_data) return fail(e, 2, "missing fmt/data chunk: " + e.audio_path);
    if (e.format != 1 || e.bits != 16) return fail(e, 3, "not 16-bit PCM: " + e.audio_path);
}

// ---- TextGrid ----
struct Token {
    bool is_string;
    std::string text;
    double number;
};

void tokenize(const std::string& s, std::vector<Token>& tokens) {
    size_t i = 0, n = s.size();
    while (i < n) {
        char c = s[i];
        if (c == '"') {
            std::string text;
            ++i;
            while (i < n) {
                if (s[i] == '"') {
                    if (i + 1 < n && s[i + 1] == '"') { text.push_back('"'); i += 2; continue; }
                    break;
                }
                text.push_back(s[i++]);
            }
            ++i;
            tokens.push_back({true, text, 0.0});
        } else if (c == '[') {                       // "intervals [3]:" index
            while (i < n && s[i] != ']') ++i;
            ++i;
        } else if ((c >= '0' && c <= '9') || ((c == '-' || c == '+') && i + 1 < n && s[i + 1] >= '0' && s[i + 1] <= '9')) {
            char* end = nullptr;
            double v = std::strtod(s.c_str() + i, &end);
            size_t used = end - (s.c_str() + i);
            if (used == 0) { ++i; continue; }
            tokens.push_back({false, std::string(), v});
            i += used;
        } else {
            ++i;
        }
    }
}

void parse_textgrid(const std::string& path, FileEntry& e) {
    std::string raw;
    if (!read_file(path, raw)) return fail(e, 4, "cannot open " + path);
    if (raw.size() >= 2 && ((unsigned char)raw[0] == 0xFF || (unsigned char)raw[0] == 0xFE) &&
        ((unsigned char)raw[1] == 0xFE || (unsigned char)raw[1] == 0xFF))
        return fail(e, 5, "UTF-16 TextGrid: " + path);
    if (raw.size() >= 3 && (unsigned char)raw[0] == 0xEF) raw.erase(0, 3);   // UTF-8 BOM
    std::vector<Token> tokens;
    tokenize(raw, tokens);
    std::vector<Word> first, chosen;
    bool have_first = false, have_chosen = false;
    for (size_t i = 0; i + 4 < tokens.size(); ++i) {
        if (!tokens[i].is_string) continue;
        const bool interval = tokens[i].text == "IntervalTier";
        if (!interval && tokens[i].text != "TextTier") continue;
        if (!tokens[i + 1].is_string || tokens[i + 4].is_string) continue;
        const std::string name = tokens[i + 1].text;
        const long count = (long)tokens[i + 4].number;
        size_t cursor = i + 5;
        if (!interval) { i = cursor + 2 * count - 1; continue; }
        std::vector<Word> words;
        bool ok = true;
        for (long j = 0; j < count; ++j, cursor += 3) {
            if (cursor + 2 >= tokens.size()) { ok = false; break; }
            if (tokens[cursor].is_string || tokens[cursor + 1].is_string || !tokens[cursor + 2].is_string) { ok = false; break; }
            words.push_back({tokens[cursor + 2].text, tokens[cursor].number, tokens[cursor + 1].number});
        }
        if (!ok) return fail(e, 6, "malformed TextGrid: " + path);
        std::string lower = name;
        for (auto& ch : lower) ch = (char)std::tolower((unsigned char)ch);
        if (!have_first) { first = words; have_first = true; }
        if (!have_chosen && (lower == "words" || lower == "word")) { chosen = words; have_chosen = true; }
        i = cursor - 1;
    }
    if (!have_first) return fail(e, 6, "no interval tier: " + path);
    const std::vector<Word>& source = have_chosen ? chosen : first;
    // fill gaps with silences so the words tile [0, end] (alignment.py:_fill_gaps)
    double at = 0.0;
    for (const Word& w : source) {
        std::string label = w.label;
        size_t a = label.find_first_not_of(" \t\r\n"), b = label.find_last_not_of(" \t\r\n");
        std::string trimmed = a == std::string::npos ? "" : label.substr(a, b - a + 1);
        if (trimmed.empty() || trimmed == "sp" || trimmed == "sil") label = kSilence;
        if (w.start - at > 1e-9) e.words.push_back({kSilence, at, w.start});
        e.words.push_back({label, w.start, w.end});
        at = w.end;
    }
}

std::string format_double(double v) {
    char buffer[64];
    auto result = std::to_chars(buffer, buffer + sizeof(buffer), v);
    std::string s(buffer, result.ptr);
    if (s.find_first_of(".en") == std::string::npos) s += ".0";
    return s;
}

bool write_textgrid(const std::string& path, const std::vector<Word>& words) {
    std::string out;
    const std::string xmax = format_double(words.empty() ? 0.0 : words.back().end);
    out += "File type = \"ooTextFile\"\nObject class = \"TextGrid\"\n\nxmin = 0\nxmax = " + xmax +
           "\ntiers? <exists>\nsize = 2\nitem []:\n";
    const char* names[2] = {"words", "phones"};
    for (int tier = 0; tier < 2; ++tier) {
        out += "    item [" + std::to_string(tier + 1) + "]:\n        class = \"IntervalTier\"\n        name = \"" +
               names[tier] + "\"\n        xmin = 0\n        xmax = " + xmax +
               "\n        intervals: size = " + std::to_string(words.size()) + "\n";
        for (size_t j = 0; j < words.size(); ++j) {
            std::string label;
            if (words[j].label == kSilence) {
                label = tier == 0 ? "sp" : "sil";
            } else {
                for (char c : words[j].label) { if (c == '"') label += "\"\""; else label.push_back(c); }
            }
            out += "        intervals [" + std::to_string(j + 1) + "]:\n            xmin = " +
                   format_double(words[j].start) + "\n            xmax = " + format_double(words[j].end) +
                   "\n            text = \"" + label + "\"\n";
        }
    }
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    size_t wrote = std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
    return wrote == out.size();
}

// ---------------------------------------------------------------------------
// torch.save(scores, f'{prefix}.pt') (emphases/core.py:112,177) without the
// Python pickler: a (1, W) float32 tensor as the zip archive torch.load reads
// (records <stem>/data.pkl, byteorder, data/0 (64-byte aligned), version; all
// stored, CRC-32 in the headers).
uint32_t crc32_of(const unsigned char* data, size_t size) {
    static uint32_t table[256];
    static std::atomic<bool> ready(false);
    if (!ready.load(std::memory_order_acquire)) {
        uint32_t local[256];
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            local[i] = c;
        }
        std::memcpy(table, local, sizeof(table));     // idempotent if two threads race
        ready.store(true, std::memory_order_release);
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < size; ++i) c = table[(c ^ data[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

void put16(std::string& out, uint32_t v) {
    out.push_back((char)(v & 0xFF));
    out.push_back((char)((v >> 8) & 0xFF));
}
void put32(std::string& out, uint32_t v) {
    put16(out, v & 0xFFFF);
    put16(out, v >> 16);
}
// pickle protocol-2 integer (BININT1 / BININT2 / BININT)
void put_pickle_int(std::string& out, uint32_t v) {
    if (v < 256) {
        out.push_back('K');
        out.push_back((char)v);
    } else if (v < 65536) {
        out.push_back('M');
        put16(out, v);
    } else {
        out.push_back('J');
        put32(out, v);
    }
}

bool write_score_file(const char* path, const float* values, uint32_t count) {
    // archive name = file stem, as torch.save chooses it
    std::string stem(path);
    size_t slash = stem.find_last_of('/');
    if (slash != std::string::npos) stem = stem.substr(slash + 1);
    size_t dot = stem.find_last_of('.');
    if (dot != std::string::npos && dot > 0) stem = stem.substr(0, dot);
    if (stem.empty()) stem = "archive";

    // pickle protocol 2 of torch._utils._rebuild_tensor_v2(FloatStorage '0' on cpu
    // with W elements, offset 0, size (1, W), stride (W, 1), requires_grad False,
    // OrderedDict()) -- byte for byte what torch.save emits; the three integers
    // are spliced in between the constant pieces
    static const unsigned char kHead[93] = {128, 2, 99, 116, 111, 114, 99, 104, 46, 95, 117, 116, 105, 108, 115, 10, 95, 114, 101, 98, 117, 105, 108, 100, 95, 116, 101, 110, 115, 111, 114, 95, 118, 50, 10, 113, 0, 40, 40, 88, 7, 0, 0, 0, 115, 116, 111, 114, 97, 103, 101, 113, 1, 99, 116, 111, 114, 99, 104, 10, 70, 108, 111, 97, 116, 83, 116, 111, 114, 97, 103, 101, 10, 113, 2, 88, 1, 0, 0, 0, 48, 113, 3, 88, 3, 0, 0, 0, 99, 112, 117, 113, 4};
    static const unsigned char kAfterNumel[8] = {116, 113, 5, 81, 75, 0, 75, 1};
    static const unsigned char kAfterShape[3] = {134, 113, 6};
    static const unsigned char kTail[44] = {75, 1, 134, 113, 7, 137, 99, 99, 111, 108, 108, 101, 99, 116, 105, 111, 110, 115, 10, 79, 114, 100, 101, 114, 101, 100, 68, 105, 99, 116, 10, 113, 8, 41, 82, 113, 9, 116, 113, 10, 82, 113, 11, 46};
    std::string pickle;
    pickle.append(reinterpret_cast<const char*>(kHead), sizeof(kHead));
    put_pickle_int(pickle, count);                  // storage numel
    pickle.append(reinterpret_cast<const char*>(kAfterNumel), sizeof(kAfterNumel));
    put_pickle_int(pickle, count);                  // size (1, W)
    pickle.append(reinterpret_cast<const char*>(kAfterShape), sizeof(kAfterShape));
    put_pickle_int(pickle, count);                  // stride (W, 1)
    pickle.append(reinterpret_cast<const char*>(kTail), sizeof(kTail));

    struct Record { std::string name; const unsigned char* data; size_t size; bool align; };
    const char* order = "little";
    const char* version = "3\n";
    const Record records[4] = {
        {stem + "/data.pkl", reinterpret_cast<const unsigned char*>(pickle.data()), pickle.size(), false},
        {stem + "/byteorder", reinterpret_cast<const unsigned char*>(order), 6, false},
        {stem + "/data/0", reinterpret_cast<const unsigned char*>(values), (size_t)count * 4, true},
        {stem + "/version", reinterpret_cast<const unsigned char*>(version), 2, false}};
    std::string out, central;
    out.reserve(512 + (size_t)count * 4 + 4 * stem.size());
    for (const Record& r : records) {
        std::string extra;
        if (r.align) {
            const size_t start = out.size() + 30 + r.name.size() + 4;
            const size_t pad = (64 - start % 64) % 64;
            extra = "FB";
            put16(extra, (uint32_t)pad);
            extra.append(pad, 'Z');
        }
        const uint32_t crc = crc32_of(r.data, r.size);
        const uint32_t offset = (uint32_t)out.size();
        put32(out, 0x04034b50u); put16(out, 20); put16(out, 0); put16(out, 0);
        put16(out, 0); put16(out, 0x21); put32(out, crc);
        put32(out, (uint32_t)r.size); put32(out, (uint32_t)r.size);
        put16(out, (uint32_t)r.name.size()); put16(out, (uint32_t)extra.size());
        out += r.name; out += extra;
        out.append(reinterpret_cast<const char*>(r.data), r.size);
        put32(central, 0x02014b50u); put16(central, 20); put16(central, 20); put16(central, 0);
        put16(central, 0); put16(central, 0); put16(central, 0x21); put32(central, crc);
        put32(central, (uint32_t)r.size); put32(central, (uint32_t)r.size);
        put16(central, (uint32_t)r.name.size()); put16(central, 0); put16(central, 0);
        put16(central, 0); put16(central, 0); put32(central, 0); put32(central, offset);
        central += r.name;
    }
    const uint32_t directory = (uint32_t)out.size();
    out += central;
    put32(out, 0x06054b50u); put16(out, 0); put16(out, 0); put16(out, 4); put16(out, 4);
    put32(out, (uint32_t)central.size()); put32(out, directory); put16(out, 0);

    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    return (std::fclose(f) == 0) && ok;
}

// A process-wide pool of worker threads: the corpus calls arrive in bursts of
// hundreds (a fill per group of files, writers per launch), and spawning 24
// threads for each costs about a millisecond.  One job runs at a time; callers
// are serialised on `submit`.
class WorkerPool {
public:
    static WorkerPool& instance() {
        static WorkerPool pool;
        return pool;
    }
    // run body(i) for i in [0, n) on up to n_threads workers (the caller helps)
    void run(int n, int n_threads, const std::function<void(int)>& body) {
        if (n <= 0) return;
        std::lock_guard<std::mutex> submit_lock(submit_);
        if (n_threads < 1) n_threads = 1;
        if (n_threads > n) n_threads = n;
        grow(n_threads - 1);
        {
            std::lock_guard<std::mutex> lock(mutex_);
            body_ = &body;
            count_ = n;
            next_.store(0);
            wanted_ = n_threads - 1;
            started_ = 0;
            pending_ = n_threads - 1;
            ++generation_;
        }
        wake_.notify_all();
        for (int i = next_.fetch_add(1); i < n; i = next_.fetch_add(1)) body(i);
        std::unique_lock<std::mutex> lock(mutex_);
        done_.wait(lock, [&] { return pending_ == 0; });
        body_ = nullptr;
    }

private:
    WorkerPool() = default;
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            stop_ = true;
            ++generation_;
        }
        wake_.notify_all();
        for (auto& thread : threads_) thread.join();
    }
    void grow(int workers) {
        while ((int)threads_.size() < workers) {
            const int index = (int)threads_.size();
            threads_.emplace_back([this, index] { loop(index); });
        }
    }
    void loop(int index) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)>* body;
            int count;
            {
                std::unique_lock<std::mutex> lock(mutex_);
                wake_.wait(lock, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                if (started_ >= wanted_) continue;     // this job needs fewer workers
                ++started_;
                body = body_;
                count = count_;
            }
            for (int i = next_.fetch_add(1); i < count; i = next_.fetch_add(1)) (*body)(i);
            {
                std::lock_guard<std::mutex> lock(mutex_);
                --pending_;
            }
            done_.notify_all();
        }
    }
    std::mutex submit_, mutex_;
    std::condition_variable wake_, done_;
    std::vector<std::thread> threads_;
    const std::function<void(int)>* body_ = nullptr;
    std::atomic<int> next_{0};
    int count_ = 0, wanted_ = 0, started_ = 0, pending_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

// many short calls (a fill per group of files, a writer per launch): the pool
template <typename Fn>
void pooled_for(int n, int n_threads, Fn fn) {
    const std::function<void(int)> body = fn;
    WorkerPool::instance().run(n, n_threads, body);
}

// one long call (it may run beside a pooled one): its own threads
template <typename Fn>
void parallel_for(int n, int n_threads, Fn fn) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n) n_threads = n > 0 ? n : 1;
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&]() {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto& t : pool) t.join();
}

}  // namespace

struct emph_corpus {
    std::vector<FileEntry> files;
};

extern "C" {

emph_corpus* emph_corpus_open(
    const char* const* text_paths, const char* const* audio_paths, int32_t n_files, int32_t n_threads) {
    emph_corpus* corpus = new emph_corpus();
    corpus->files.resize(n_files > 0 ? n_files : 0);
    parallel_for(n_files, n_threads, [&](int i) {
        FileEntry& e = corpus->files[i];
        e.audio_path = audio_paths[i];
        scan_wav(e);
        parse_textgrid(text_paths[i], e);
    });
    return corpus;
}

int emph_corpus_info(
    const emph_corpus* corpus, int32_t* status, int32_t* sample_rate, int32_t* channels,
    int64_t* n_samples, int32_t* n_words) {
    if (!corpus) return EMPH_EINVAL;
    for (size_t i = 0; i < corpus->files.size(); ++i) {
        const FileEntry& e = corpus->files[i];
        status[i] = e.status;
        sample_rate[i] = e.sample_rate;
        channels[i] = e.channels;
        n_samples[i] = e.n_samples;
        n_words[i] = (int32_t)e.words.size();
    }
    return EMPH_OK;
}

const char* emph_corpus_error(const emph_corpus* corpus, int32_t index) {
    if (!corpus || index < 0 || (size_t)index >= corpus->files.size()) return "";
    return corpus->files[index].error.c_str();
}

int emph_corpus_fill_files(
    emph_corpus* corpus, const int32_t* file_indices, int32_t n_indices,
    int16_t* audio_dst, const int64_t* sample_offsets,
    double* times_dst, const int64_t* word_offsets, int32_t n_threads) {
    if (!corpus || n_indices < 0 || (n_indices > 0 && !file_indices)) return EMPH_EINVAL;
    const int n_files = (int)corpus->files.size();
    std::atomic<int> failures(0);
    pooled_for(n_indices, n_threads, [&](int k) {
        const int i = file_indices[k];
        if (i < 0 || i >= n_files) { ++failures; return; }
        FileEntry& e = corpus->files[i];
        if (e.status != 0) return;
        double* times = times_dst + 2 * word_offsets[i];
        for (size_t j = 0; j < e.words.size(); ++j) {
            times[2 * j] = e.words[j].start;
            times[2 * j + 1] = e.words[j].end;
        }
        FILE* f = std::fopen(e.audio_path.c_str(), "rb");
        if (!f) { fail(e, 1, "cannot reopen " + e.audio_path); ++failures; return; }
        std::fseek(f, (long)e.data_offset, SEEK_SET);
        int16_t* dst = audio_dst + sample_offsets[i];
        bool ok = true;
        if (e.channels == 1) {
            ok = std::fread(dst, 2, (size_t)e.n_samples, f) == (size_t)e.n_samples;
        } else {                                     // keep channel 0 (mels.py:48)
            std::vector<int16_t> frame_buffer((size_t)e.channels * 4096);
            long long done = 0;
            while (done < e.n_samples && ok) {
                long long want = e.n_samples - done < 4096 ? e.n_samples - done : 4096;
                ok = std::fread(frame_buffer.data(), 2 * e.channels, (size_t)want, f) == (size_t)want;
                for (long long k2 = 0; k2 < want; ++k2) dst[done + k2] = frame_buffer[(size_t)k2 * e.channels];
                done += want;
            }
        }
        std::fclose(f);
        if (!ok) { fail(e, 1, "short read: " + e.audio_path); ++failures; }
    });
    return failures.load() == 0 ? EMPH_OK : EMPH_EINVAL;
}

int emph_corpus_fill(
    emph_corpus* corpus, int16_t* audio_dst, const int64_t* sample_offsets,
    double* times_dst, const int64_t* word_offsets, int32_t n_threads) {
    if (!corpus) return EMPH_EINVAL;
    std::vector<int32_t> all(corpus->files.size());
    for (size_t i = 0; i < all.size(); ++i) all[i] = (int32_t)i;
    return emph_corpus_fill_files(
        corpus, all.data(), (int32_t)all.size(), audio_dst, sample_offsets, times_dst,
        word_offsets, n_threads);
}

int emph_corpus_write_textgrids(
    const emph_corpus* corpus, const char* const* output_paths, int32_t n_threads) {
    if (!corpus) return EMPH_EINVAL;
    std::atomic<int> failures(0);
    parallel_for((int)corpus->files.size(), n_threads, [&](int i) {
        const FileEntry& e = corpus->files[i];
        if (e.status != 0 || output_paths[i] == nullptr || output_paths[i][0] == 0) return;
        if (!write_textgrid(output_paths[i], e.words)) ++failures;
    });
    return failures.load() == 0 ? EMPH_OK : EMPH_EINVAL;
}

int emph_pack_audio_f32(
    const float* const* sources, const int64_t* lengths, const int64_t* offsets,
    int32_t n_utterances, float* dst_f32, int16_t* dst_i16, int32_t* narrowed,
    int32_t n_threads) {
    if (n_utterances < 0 || (n_utterances > 0 && (!sources || !lengths || !offsets || !dst_f32)))
        return EMPH_EINVAL;
    if (narrowed) *narrowed = 0;
    // Lossless narrowing: decoded 16-bit PCM (what load.audio returns) is
    // k / 32768 with integer k in [-32768, 32767]; then the int16 copy holds the
    // same values in half the bytes.  A probe of each utterance's head decides
    // before the full pass; the full pass re-checks every sample.
    auto exact = [](float v, int16_t& out) {
        const float scaled = v * 32768.f;
        if (!(scaled >= -32768.f && scaled <= 32767.f)) return false;   // also NaN
        const int k = (int)scaled;
        out = (int16_t)k;
        return (float)k == scaled;
    };
    bool try_narrow = dst_i16 != nullptr && narrowed != nullptr;
    if (try_narrow) {
        for (int i = 0; i < n_utterances && try_narrow; ++i) {
            const int64_t probe = lengths[i] < 64 ? lengths[i] : 64;
            int16_t unused;
            for (int64_t k = 0; k < probe; ++k)
                if (!exact(sources[i][k], unused)) { try_narrow = false; break; }
        }
    }
    if (try_narrow) {
        std::atomic<int> inexact(0);
        parallel_for(n_utterances, n_threads, [&](int i) {
            const float* src = sources[i];
            int16_t* dst = dst_i16 + offsets[i];
            const int64_t n = lengths[i];
            for (int64_t base = 0; base < n && !inexact.load(std::memory_order_relaxed); base += 4096) {
                const int64_t end = base + 4096 < n ? base + 4096 : n;
                int bad = 0;
                int64_t k = base;
#if defined(__SSE2__)
                // 8 samples per step: scale, truncate, compare back, saturating
                // pack (a value outside int16 fails the range compare)
                const __m128 scale = _mm_set1_ps(32768.f);
                const __m128 low = _mm_set1_ps(-32768.f), high = _mm_set1_ps(32767.f);
                __m128 wrong = _mm_setzero_ps();
                for (; k + 8 <= end; k += 8) {
                    const __m128 a = _mm_mul_ps(_mm_loadu_ps(src + k), scale);
                    const __m128 b = _mm_mul_ps(_mm_loadu_ps(src + k + 4), scale);
                    const __m128i ia = _mm_cvttps_epi32(a), ib = _mm_cvttps_epi32(b);
                    wrong = _mm_or_ps(wrong, _mm_cmpneq_ps(_mm_cvtepi32_ps(ia), a));
                    wrong = _mm_or_ps(wrong, _mm_cmpneq_ps(_mm_cvtepi32_ps(ib), b));
                    wrong = _mm_or_ps(wrong, _mm_or_ps(_mm_cmplt_ps(a, low), _mm_cmpgt_ps(a, high)));
                    wrong = _mm_or_ps(wrong, _mm_or_ps(_mm_cmplt_ps(b, low), _mm_cmpgt_ps(b, high)));
                    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + k), _mm_packs_epi32(ia, ib));
                }
                bad |= _mm_movemask_ps(wrong);
#endif
                for (; k < end; ++k) bad |= !exact(src[k], dst[k]);
                if (bad) inexact.store(1, std::memory_order_relaxed);
            }
        });
        if (!inexact.load()) {
            *narrowed = 1;
            return EMPH_OK;
        }
    }
    parallel_for(n_utterances, n_threads, [&](int i) {
        std::memcpy(dst_f32 + offsets[i], sources[i], (size_t)lengths[i] * sizeof(float));
    });
    return EMPH_OK;
}

int emph_write_score_files(
    const char* const* paths, const float* scores, const int64_t* offsets,
    const int32_t* counts, int32_t n_files, int32_t n_threads) {
    if (n_files < 0 || (n_files > 0 && (!paths || !scores || !offsets || !counts))) return EMPH_EINVAL;
    std::atomic<int> failures(0);
    parallel_for(n_files, n_threads, [&](int i) {
        if (paths[i] == nullptr || paths[i][0] == 0) return;
        if (counts[i] < 0 || !write_score_file(paths[i], scores + offsets[i], (uint32_t)counts[i]))
            ++failures;
    });
    return failures.load() == 0 ? EMPH_OK : EMPH_EINVAL;
}

int emph_write_score_rows(
    const char* const* paths, const float* const* rows, const int32_t* counts,
    int32_t n_files, int32_t n_threads) {
    if (n_files < 0 || (n_files > 0 && (!paths || !rows || !counts))) return EMPH_EINVAL;
    std::atomic<int> failures(0);
    auto write = [&](int i) {
        if (paths[i] == nullptr || paths[i][0] == 0) return;
        if (counts[i] < 0 || (counts[i] > 0 && !rows[i]) ||
            !write_score_file(paths[i], rows[i], (uint32_t)counts[i]))
            ++failures;
    };
    // n_threads < 0: -n_threads threads of this call's own (it runs beside a
    // decode that occupies the shared pool)
    if (n_threads < 0) parallel_for(n_files, -n_threads, write);
    else pooled_for(n_files, n_threads, write);
    return failures.load() == 0 ? EMPH_OK : EMPH_EINVAL;
}

void emph_corpus_close(emph_corpus* corpus) { delete corpus; }

// ---- the same entry points with the paths as ONE buffer of NUL-terminated
// strings (n of them, back to back): building a char*[] of tens of thousands
// of Python strings costs more than parsing the files ----
static std::vector<const char*> split_blob(const char* blob, int32_t n) {
    std::vector<const char*> out((size_t)(n > 0 ? n : 0));
    const char* cursor = blob;
    for (int32_t i = 0; i < n; ++i) {
        out[i] = cursor;
        cursor += std::strlen(cursor) + 1;
    }
    return out;
}

emph_corpus* emph_corpus_open_blob(
    const char* text_blob, const char* audio_blob, int32_t n_files, int32_t n_threads) {
    const std::vector<const char*> text = split_blob(text_blob, n_files);
    const std::vector<const char*> audio = split_blob(audio_blob, n_files);
    return emph_corpus_open(text.data(), audio.data(), n_files, n_threads);
}

/* paths of the files with mask[i] != 0 only, in file order */
int emph_corpus_write_textgrids_blob(
    const emph_corpus* corpus, const char* path_blob, const uint8_t* mask, int32_t n_threads) {
    if (!corpus || !mask) return EMPH_EINVAL;
    const int32_t n = (int32_t)corpus->files.size();
    std::vector<const char*> paths((size_t)n, nullptr);
    const char* cursor = path_blob;
    for (int32_t i = 0; i < n; ++i) {
        if (!mask[i]) continue;
        paths[i] = cursor;
        cursor += std::strlen(cursor) + 1;
    }
    return emph_corpus_write_textgrids(corpus, paths.data(), n_threads);
}

/* file i holds counts[i] values starting at base[starts[i]] */
int emph_write_score_rows_blob(
    const char* path_blob, const float* base, const int64_t* starts, const int32_t* counts,
    int32_t n_files, int32_t n_threads) {
    if (n_files < 0 || (n_files > 0 && (!path_blob || !base || !starts || !counts))) return EMPH_EINVAL;
    const std::vector<const char*> paths = split_blob(path_blob, n_files);
    std::vector<const float*> rows((size_t)n_files);
    for (int32_t i = 0; i < n_files; ++i) rows[i] = base + starts[i];
    return emph_write_score_rows(paths.data(), rows.data(), counts, n_files, n_threads);
}

/* sizes[i] = size in bytes of file i of a NUL-separated path buffer (-1: cannot
 * stat): the cost proxy of the length-balanced sharding, without 24,000
 * Python-level stat calls on every rank */
int emph_file_sizes(const char* path_blob, int32_t n_files, int32_t n_threads, int64_t* sizes) {
    if (n_files < 0 || (n_files > 0 && (!path_blob || !sizes))) return EMPH_EINVAL;
    const std::vector<const char*> paths = split_blob(path_blob, n_files);
    pooled_for(n_files, n_threads, [&](int i) {
        struct stat info;
        sizes[i] = stat(paths[i], &info) == 0 ? (int64_t)info.st_size : -1;
    });
    return EMPH_OK;
}

}  // extern "C"
