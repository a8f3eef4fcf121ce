// ntstats.cu -- base composition and contig structure of a bin's scaffolds: the integer half of CheckM's bin statistics
// (checkm/binStatistics.py:176-243: calculateGC, calculateSeqStats).  Everything here is a byte scan
// bound by HBM: a CTA walks a 128 KB segment of a scaffold in 16 KB tiles, a thread takes 64 consecutive bytes.
//
// What the reference computes per scaffold, restated as local predicates:
//   * a, c, g, t: case-insensitive counts, U counted with T (util/seqUtils.py:279-286)
//   * ambiguous bases: count('N') + count('n')
//   * contigs: scaffold.split('N' * 10), the remaining 'N' of every piece deleted, empty pieces dropped.  A maximal run of
//     r upper-case N holds floor(r / 10) separators and the other N are deleted anyway, so the pieces are exactly the
//     stretches between runs of >= 10 N, and a piece's length is its number of bytes that are not 'N'.
//     Position i ends such a run iff bytes i-9..i are all 'N' and byte i+1 is not (or the scaffold ends): a predicate with a
//     9-byte halo.  The contig index of a byte is the number of run ends before it (a prefix sum), the contig length a
//     histogram over that index.
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>
#include "engine.hpp"
#include "pool.hpp"

using namespace ckm;

namespace {

constexpr int NT_THREADS = 256;
constexpr int NT_CHUNK = 64;                         // bytes of one thread in one tile: one bit each in a 64-bit mask
constexpr int NT_TILE = NT_THREADS * NT_CHUNK;       // 16 KB
constexpr int NT_SEG_TILES = 8;
constexpr int64_t NT_SEG = (int64_t)NT_TILE * NT_SEG_TILES;   // 128 KB: the unit of work of a CTA
constexpr int NT_MAXSEG = NT_TILE / 11 + 8;          // run ends are >= 11 bytes apart
constexpr int NT_WARPS = NT_THREADS / 32;

// A scaffold longer than NT_SEG is cut into segments scanned by different CTAs.  Contigs closed inside a segment are
// reported by the kernel; of each segment the bases before its first run end (head) and after its last (tail) come back
// separately and the host joins tail + head across the cuts (join_segments below).
struct NtSegment { uint32_t head, tail, closed; };   // closed: the segment holds at least one run end

struct NtParams {
  const uint8_t *bytes;            // every scaffold starts at a multiple of 64 and is followed by padding up to the next one
  const int64_t *starts, *lens;
  const int32_t *seg_scaf;         // segments in scaffold order
  const int64_t *seg_off;          // first byte of the segment inside its scaffold (a multiple of NT_SEG)
  int32_t nseg;
  int32_t *work;                   // next segment
  NtSegment *seg;
  unsigned long long *stats;       // nscaf x 8: a c g t N n contigs contig_bases (the last two: contigs closed inside segments)
  uint32_t *contig_scaf, *contig_len;
  unsigned long long *ncontigs;
  long long cap;
};

// 0x80 in every byte of w that equals the byte replicated in pat
__device__ __forceinline__ uint32_t eq4(uint32_t w, uint32_t pat) {
  const uint32_t x = w ^ pat;
  const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  return ~(t | x | 0x7F7F7F7Fu);
}
// the four 0x80 flags of eq4 as a 4-bit mask, byte 0 -> bit 0
__device__ __forceinline__ uint32_t nibble(uint32_t flags) { return (((flags >> 7) * 0x01020408u) >> 24) & 0xFu; }
// sum of the four bytes of x (each below 64)
__device__ __forceinline__ uint32_t hsum4(uint32_t x) { return (x * 0x01010101u) >> 24; }

__global__ void __launch_bounds__(NT_THREADS, 4) ntstats_kernel(NtParams p) {
  __shared__ unsigned long long s_m[NT_THREADS];
  __shared__ uint32_t s_acc[NT_MAXSEG];
  __shared__ uint32_t s_wsum[NT_WARPS];
  __shared__ unsigned long long s_tot[8];
  __shared__ int s_next;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NT_MAXSEG; i += NT_THREADS) s_acc[i] = 0;
  if (tid < 8) s_tot[tid] = 0;
  __syncthreads();
  for (;;) {
    if (tid == 0) s_next = atomicAdd(p.work, 1);
    __syncthreads();
    const int item = s_next;
    if (item >= p.nseg) return;
    const int s = p.seg_scaf[item];
    const int64_t L = p.lens[s];
    const int64_t seg_begin = p.seg_off[item];
    const int64_t seg_end = min(L, seg_begin + NT_SEG);
    const uint8_t *base = p.bytes + p.starts[s];
    uint32_t cA = 0, cC = 0, cG = 0, cT = 0, cN = 0, cn = 0;     // per-thread counts over the segment
    uint32_t carry = 0;                                          // bases of the contig still open (same in every thread)
    uint32_t head = 0; bool closed = false;                      // same in every thread
    uint32_t my_ctg = 0; unsigned long long my_ctg_bases = 0;
    for (int64_t t0 = seg_begin; t0 < seg_end; t0 += NT_TILE) {
      const int64_t pos = t0 + (int64_t)tid * NT_CHUNK;
      unsigned long long m = 0, valid = 0;
      if (pos < L) {
        const int64_t left = L - pos;
        valid = left >= NT_CHUNK ? ~0ull : ((1ull << left) - 1ull);
        const uint4 *src = reinterpret_cast<const uint4 *>(base + pos);
        uint32_t w[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const uint4 v = __ldg(src + q); w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
        // Fast path, a chunk of nothing but upper-case A C G T (what assemblies mostly are): the low three bits of the
        // four letters differ (A 1, C 3, T 4, G 7), so one byte permute looks up the letter each byte would have to be and
        // one xor tells whether it is; then bits 1 and 2 of the byte give the letter: A 00, C 01, T 10, G 11.
        uint32_t bad = left >= NT_CHUNK ? 0u : 1u;
        uint32_t s1 = 0, s2 = 0, sg = 0;                         // per-byte sums over the 16 words: bit1, bit2, bit1 & bit2
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint32_t t = w[j] & 0x07070707u;
          t |= t >> 4;
          const uint32_t sel = __byte_perm(t, 0u, 0x4420);       // the four 3-bit indices as selector nibbles
          const uint32_t expect = __byte_perm(0x43FF41FFu, 0x47FFFF54u, sel);
          bad |= w[j] ^ expect;
          const uint32_t b1 = w[j] >> 1, b2 = w[j] >> 2;
          s1 += b1 & 0x01010101u; s2 += b2 & 0x01010101u; sg += b1 & b2 & 0x01010101u;
        }
        if (bad == 0) {
          const uint32_t n1 = hsum4(s1), n2 = hsum4(s2), ng = hsum4(sg);
          cG += ng; cC += n1 - ng; cT += n2 - ng; cA += NT_CHUNK - n1 - n2 + ng;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int first = j * 4;
            uint32_t x = w[j];
            if (left < first + 4) {               // last chunk of the scaffold: whatever the padding holds is not sequence
              const int keep = (int)(left - first);
              x = keep <= 0 ? 0u : (x & ((1u << (8 * keep)) - 1u));
            }
            const uint32_t up = x & 0xDFDFDFDFu;             // 'a' -> 'A'; no other byte maps onto a letter tested below
            cA += __popc(eq4(up, 0x41414141u));
            cC += __popc(eq4(up, 0x43434343u));
            cG += __popc(eq4(up, 0x47474747u));
            cT += __popc(eq4(up, 0x54545454u)) + __popc(eq4(up, 0x55555555u));
            const uint32_t fN = eq4(x, 0x4E4E4E4Eu);
            cN += __popc(fN);
            cn += __popc(eq4(x, 0x6E6E6E6Eu));
            m |= (unsigned long long)nibble(fN) << first;
          }
        }
      }
      s_m[tid] = m;
      __syncthreads();
      // halo: is-N of the 9 bytes before this chunk and of the byte after it
      unsigned long long prev9 = 0, nextbit = 0;
      if (tid > 0) prev9 = s_m[tid - 1] >> 55;
      else if (pos > 0 && pos < L) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(base + pos - 16));
        const uint32_t bits = nibble(eq4(v.y, 0x4E4E4E4Eu)) | (nibble(eq4(v.z, 0x4E4E4E4Eu)) << 4) | (nibble(eq4(v.w, 0x4E4E4E4Eu)) << 8);
        prev9 = bits >> 3;                                   // bytes pos-12..pos-1 -> the last nine
      }
      if (tid < NT_THREADS - 1) nextbit = s_m[tid + 1] & 1ull;
      else if (pos + NT_CHUNK < L) nextbit = base[pos + NT_CHUNK] == 'N';
      unsigned long long ends = 0;
      if (m | prev9) {
        const unsigned __int128 X = ((unsigned __int128)m << 9) | (unsigned __int128)prev9;
        const unsigned __int128 A = X & (X >> 1), B = A & (A >> 2), C8 = B & (B >> 4);
        const unsigned long long run10 = (unsigned long long)(C8 & (A >> 8));   // bit k: bytes pos+k-9 .. pos+k are all N
        ends = run10 & ~((m >> 1) | (nextbit << 63)) & valid;
      }
      const unsigned long long bases = valid & ~m;
      // contig index of the chunk's first byte, relative to the tile: exclusive scan of the run ends
      const uint32_t nb = __popcll(ends);
      uint32_t incl = nb;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      uint32_t before = 0, total = 0;
#pragma unroll
      for (int w8 = 0; w8 < NT_WARPS; ++w8) { const uint32_t v = s_wsum[w8]; if (w8 < warp) before += v; total += v; }
      uint32_t id = before + incl - nb;
      if (__all_sync(0xffffffffu, nb == 0)) {                // the whole warp lies inside one contig
        uint32_t c = __popcll(bases);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if (lane == 0 && c) atomicAdd(&s_acc[id], c);
      } else {
        unsigned long long rest = bases, e = ends;
        while (e) {
          const int b = __ffsll((long long)e) - 1;
          const unsigned long long upto = b == 63 ? ~0ull : ((2ull << b) - 1ull);
          const uint32_t c = __popcll(rest & upto);
          if (c) atomicAdd(&s_acc[id], c);
          rest &= ~upto; e &= e - 1; ++id;
        }
        const uint32_t c = __popcll(rest);
        if (c) atomicAdd(&s_acc[id], c);
      }
      __syncthreads();
      // contigs 0 .. total-1 of this tile are closed; the last index stays open into the next tile.  The first contig the
      // segment closes may have begun in the segment before: it is the segment's head, joined on the host.
      if (total > 0 && !closed) head = s_acc[0] + carry;
      for (uint32_t j = tid; j < total; j += NT_THREADS) {
        if (j == 0 && !closed) continue;
        const uint32_t len = s_acc[j] + (j == 0 ? carry : 0u);
        if (len) {
          const unsigned long long at = atomicAdd(p.ncontigs, 1ull);
          if ((long long)at < p.cap) { p.contig_scaf[at] = (uint32_t)s; p.contig_len[at] = len; }
          ++my_ctg; my_ctg_bases += len;
        }
      }
      carry = s_acc[total] + (total == 0 ? carry : 0u);
      closed = closed || total > 0;
      __syncthreads();
      for (uint32_t j = tid; j <= total; j += NT_THREADS) s_acc[j] = 0;
      // the next tile's first barrier orders these stores before its atomics
    }
    if (tid == 0) { NtSegment r; r.head = head; r.tail = carry; r.closed = closed ? 1u : 0u; p.seg[item] = r; }
    unsigned long long v[8] = {cA, cC, cG, cT, cN, cn, my_ctg, my_ctg_bases};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned long long x = v[k];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
      if (lane == 0 && x) atomicAdd(&s_tot[k], x);
    }
    __syncthreads();
    if (tid < 8) { if (s_tot[tid]) atomicAdd(&p.stats[(size_t)s * 8 + tid], s_tot[tid]); s_tot[tid] = 0; }
    // s_next is rewritten by thread 0 only after the barrier at the top of the loop, which also orders the s_tot reset
    __syncthreads();
  }
}

}  // namespace

extern "C" {

// util/seqUtils.py:180-211 readFasta, as Python's text mode feeds it lines: "\n", "\r\n" and "\r" all end a line; lines that
// are blank after strip() are skipped; a header starts a record; of every other line the LAST CHARACTER IS DROPPED as its
// newline -- so a final line without one loses a base, as in the reference.
int ckm_fasta_scan_nt(const char *text, int64_t n, uint8_t *bytes_out, int64_t bytes_cap, int64_t *starts_out, int64_t *lens_out,
                      int32_t max_records, char *headers_out, int64_t headers_cap, int32_t *nrec_out, int64_t *bytes_used_out,
                      int64_t *hdr_bytes_out) {
  if ((!text && n > 0) || !bytes_out || !starts_out || !lens_out || !headers_out || !nrec_out || !bytes_used_out || !hdr_bytes_out) {
    set_error("ckm_fasta_scan_nt: bad argument"); return CKM_EINVAL;
  }
  int32_t nrec = 0; int64_t used = 0, hb = 0, i = 0;
  auto close_record = [&]() {
    if (nrec == 0) return;
    const int64_t end = starts_out[nrec - 1] + lens_out[nrec - 1];
    const int64_t padded = (end + 63) / 64 * 64;
    std::memset(bytes_out + end, 0, (size_t)(padded - end));
    used = padded;
  };
  while (i < n) {
    const char *nl = (const char *)std::memchr(text + i, '\n', (size_t)(n - i));
    int64_t e = nl ? (nl - text) : n;                           // candidate line [i, e), terminator at e (or none)
    const char *cr = (const char *)std::memchr(text + i, '\r', (size_t)(e - i));
    int64_t next = e + 1; bool terminated = nl != nullptr;
    if (cr) { e = cr - text; terminated = true; next = (e + 1 < n && text[e + 1] == '\n') ? e + 2 : e + 1; }
    bool blank = true;
    for (int64_t j = i; j < e && blank; ++j) { const unsigned char c = (unsigned char)text[j]; blank = (c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31)); }
    if (!blank) {
      if (text[i] == '>') {
        close_record();
        if (nrec >= max_records) { set_error("ckm_fasta_scan_nt: more records than the caller allowed for"); return CKM_ECAPACITY; }
        const int64_t len = e - (i + 1);
        if (hb + len + 1 > headers_cap) { set_error("ckm_fasta_scan_nt: header buffer too small"); return CKM_ECAPACITY; }
        if (nrec > 0) headers_out[hb++] = '\n';
        std::memcpy(headers_out + hb, text + i + 1, (size_t)len); hb += len;
        starts_out[nrec] = used; lens_out[nrec] = 0; ++nrec;
      } else {
        if (nrec == 0) { set_error("ckm_fasta_scan_nt: sequence data before the first '>' line"); return CKM_EFORMAT; }
        const int64_t len = terminated ? (e - i) : (e - i - 1);
        const int64_t at = starts_out[nrec - 1] + lens_out[nrec - 1];
        if (at + len + 64 > bytes_cap) { set_error("ckm_fasta_scan_nt: output buffer too small"); return CKM_ECAPACITY; }
        std::memcpy(bytes_out + at, text + i, (size_t)len);
        lens_out[nrec - 1] += len;
      }
    }
    i = next;
  }
  close_record();
  *nrec_out = nrec; *bytes_used_out = used; *hdr_bytes_out = hb;
  return CKM_OK;
}

int ckm_scaffold_stats(ckm_engine *e, const uint8_t *bytes, int64_t nbytes, const int64_t *starts, const int64_t *lens,
                       int32_t nscaf, int64_t *stats_out, uint32_t *contig_scaffold_out, uint32_t *contig_len_out,
                       int64_t contig_cap, int64_t *ncontigs_out, float *kernel_ms_out) {
  if (!e || nscaf < 0 || nbytes < 0 || (nscaf > 0 && (!bytes || !starts || !lens || !stats_out)) || contig_cap < 0 ||
      (contig_cap > 0 && (!contig_scaffold_out || !contig_len_out)) || !ncontigs_out) {
    set_error("ckm_scaffold_stats: bad argument"); return CKM_EINVAL;
  }
  *ncontigs_out = 0;
  if (kernel_ms_out) *kernel_ms_out = 0.0f;
  if (nscaf == 0) return CKM_OK;
  for (int32_t s = 0; s < nscaf; ++s) {
    if ((starts[s] & 63) || lens[s] < 0 || lens[s] > 0xFFFFFFFFll || starts[s] < 0 || (starts[s] + lens[s] + 63) / 64 * 64 > nbytes) {
      set_error("ckm_scaffold_stats: every scaffold must start at a multiple of 64 bytes and lie, padded to 64, inside the buffer");
      return CKM_EINVAL;
    }
  }
  cudaSetDevice(e->device);
  PoolScope pool_scope(e);
  cudaStream_t st = e->stream;
  std::vector<int32_t> seg_scaf; std::vector<int64_t> seg_off;
  for (int32_t s = 0; s < nscaf; ++s)
    for (int64_t off = 0; off < lens[s]; off += NT_SEG) { seg_scaf.push_back(s); seg_off.push_back(off); }
  const int64_t nseg = (int64_t)seg_scaf.size();
  std::memset(stats_out, 0, sizeof(int64_t) * 8 * nscaf);
  if (nseg == 0) return CKM_OK;
  if (nseg > 0x7FFFFFFFll) { set_error("ckm_scaffold_stats: too many bytes for one call"); return CKM_EINVAL; }
  DevBuf dbytes, dstarts, dlens, dsegs, dsego, dseg, dstats, dcs, dcl, dctr;
  int rc;
  if ((rc = dbytes.alloc((size_t)nbytes + 64)) || (rc = dstarts.alloc(sizeof(int64_t) * nscaf)) || (rc = dlens.alloc(sizeof(int64_t) * nscaf)) ||
      (rc = dsegs.alloc(sizeof(int32_t) * nseg)) || (rc = dsego.alloc(sizeof(int64_t) * nseg)) || (rc = dseg.alloc(sizeof(NtSegment) * nseg)) ||
      (rc = dstats.alloc(sizeof(int64_t) * 8 * nscaf)) ||
      (rc = dcs.alloc(sizeof(uint32_t) * (size_t)contig_cap)) || (rc = dcl.alloc(sizeof(uint32_t) * (size_t)contig_cap)) || (rc = dctr.alloc(64)))
    return rc;
  CKM_CUDA(cudaMemcpyAsync(dbytes.p, bytes, (size_t)nbytes, cudaMemcpyHostToDevice, st));
  CKM_CUDA(cudaMemcpyAsync(dstarts.p, starts, sizeof(int64_t) * nscaf, cudaMemcpyHostToDevice, st));
  CKM_CUDA(cudaMemcpyAsync(dlens.p, lens, sizeof(int64_t) * nscaf, cudaMemcpyHostToDevice, st));
  CKM_CUDA(cudaMemcpyAsync(dsegs.p, seg_scaf.data(), sizeof(int32_t) * nseg, cudaMemcpyHostToDevice, st));
  CKM_CUDA(cudaMemcpyAsync(dsego.p, seg_off.data(), sizeof(int64_t) * nseg, cudaMemcpyHostToDevice, st));
  CKM_CUDA(cudaMemsetAsync(dctr.p, 0, 64, st));
  CKM_CUDA(cudaMemsetAsync(dstats.p, 0, sizeof(int64_t) * 8 * nscaf, st));
  NtParams p;
  p.bytes = dbytes.as<uint8_t>(); p.starts = dstarts.as<int64_t>(); p.lens = dlens.as<int64_t>();
  p.seg_scaf = dsegs.as<int32_t>(); p.seg_off = dsego.as<int64_t>(); p.nseg = (int32_t)nseg; p.seg = dseg.as<NtSegment>();
  p.work = dctr.as<int32_t>(); p.stats = dstats.as<unsigned long long>();
  p.contig_scaf = dcs.as<uint32_t>(); p.contig_len = dcl.as<uint32_t>();
  p.ncontigs = reinterpret_cast<unsigned long long *>(dctr.as<uint8_t>() + 8); p.cap = contig_cap;
  const int per_sm = 4;                                        // 4 x 256 threads at <= 64 registers
  const int grid = (int)std::min<int64_t>((int64_t)e->prop.multiProcessorCount * per_sm, nseg);
  CKM_CUDA(cudaEventRecord(e->ev[0], st));
  ntstats_kernel<<<grid, NT_THREADS, 0, st>>>(p);
  CKM_CUDA(cudaGetLastError());
  CKM_CUDA(cudaEventRecord(e->ev[1], st));
  unsigned long long n_dev = 0;
  std::vector<NtSegment> seg((size_t)nseg);
  CKM_CUDA(cudaMemcpyAsync(&n_dev, p.ncontigs, sizeof(n_dev), cudaMemcpyDeviceToHost, st));
  CKM_CUDA(cudaMemcpyAsync(stats_out, dstats.p, sizeof(int64_t) * 8 * nscaf, cudaMemcpyDeviceToHost, st));
  CKM_CUDA(cudaMemcpyAsync(seg.data(), dseg.p, sizeof(NtSegment) * nseg, cudaMemcpyDeviceToHost, st));
  CKM_CUDA(cudaStreamSynchronize(st));
  if (kernel_ms_out) CKM_CUDA(cudaEventElapsedTime(kernel_ms_out, e->ev[0], e->ev[1]));
  // join the open ends of the segments: a contig runs from the tail of one segment through every segment without a run
  // end into the head of the next one that has one
  std::vector<std::pair<uint32_t, uint32_t>> joined;
  for (int64_t i = 0; i < nseg;) {
    const int32_t s = seg_scaf[i];
    uint64_t open = 0;
    for (; i < nseg && seg_scaf[i] == s; ++i) {
      if (seg[i].closed) { if (open + seg[i].head) joined.emplace_back((uint32_t)s, (uint32_t)(open + seg[i].head)); open = seg[i].tail; }
      else open += seg[i].tail;
    }
    if (open) joined.emplace_back((uint32_t)s, (uint32_t)open);
  }
  const int64_t n_found = (int64_t)n_dev + (int64_t)joined.size();
  *ncontigs_out = n_found;
  if (n_found > contig_cap) { set_error("ckm_scaffold_stats: more contigs than the caller allowed for (the count is returned; call again)"); return CKM_ECAPACITY; }
  if (n_dev) {
    CKM_CUDA(cudaMemcpyAsync(contig_scaffold_out, dcs.p, sizeof(uint32_t) * n_dev, cudaMemcpyDeviceToHost, st));
    CKM_CUDA(cudaMemcpyAsync(contig_len_out, dcl.p, sizeof(uint32_t) * n_dev, cudaMemcpyDeviceToHost, st));
    CKM_CUDA(cudaStreamSynchronize(st));
  }
  for (size_t k = 0; k < joined.size(); ++k) {
    contig_scaffold_out[n_dev + k] = joined[k].first; contig_len_out[n_dev + k] = joined[k].second;
    stats_out[(size_t)joined[k].first * 8 + 6] += 1; stats_out[(size_t)joined[k].first * 8 + 7] += joined[k].second;
  }
  return CKM_OK;
}

}  // extern "C"
