// Package colttb200 is the cgo shim that lets coltt's Go host code drive libcoltt_b200.so.
//
// It is WRITTEN BUT NOT COMPILED in this repository's build image (no Go toolchain, no network;
// DESIGN.md §1).  It implements the search/ingest half of edge.vectorspace
// (edge/vectorstore.go:30-49) and wraps vectorindex.Hnsw's Load/Search
// (core/vectorindex/hnsw.go:243-278, hnsw_commit.go:164-278) over include/coltt_b200.h.
// Metadata maps, the inverted index and filter expressions stay on the Go side: the shim hands
// ids and vectors across the boundary and maps returned ids back to metadata.
//
//	build: CGO_CFLAGS="-I${COLTT_B200}/include" CGO_LDFLAGS="-L${COLTT_B200}/coltt_b200/lib -lcoltt_b200" go build ./...
package colttb200

/*
#cgo LDFLAGS: -lcoltt_b200
#include <stdlib.h>
#include "coltt_b200.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"
)

// Select / math modes (coltt_select, coltt_math).
const (
	SelectCompat  = 0 // the reference's literal behaviour: K largest distances (SURVEY F1)
	SelectNearest = 1
	MathExact     = 0
	MathFast      = 1
)

// call runs one C-ABI call and, on failure, fetches its message.  coltt_b200_last_error() is thread-local on the C
// side and a goroutine may migrate between OS threads from one cgo call to the next, so the call and the read of
// its message are pinned to one OS thread.
func call(f func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := f(); rc != 0 {
		return errors.New(C.GoString(C.coltt_b200_last_error()))
	}
	return nil
}

// errDim is the reference's dimension error (edge/none_vectorstore.go:86-88).
func errDim(want, got int) error {
	return fmt.Errorf("Dim Length UnmatchdError: expect dimension: [%d], but got [%d]", want, got)
}

var errTopK = errors.New("topK must be positive")

// Store is one edge collection's vectors on one GPU.
type Store struct {
	h   *C.coltt_store
	dim int
}

// NewStore replaces newNoneVectorstore / newF16Vectorstore / newBF16Vectorstore / newF8Vectorstore
// (edge/vectorstore.go:62-85).  distance and quantization are the edgepb enum values.
func NewStore(dim uint32, distance, quantization int32, device int32, capacityHint uint64) (*Store, error) {
	cfg := C.coltt_store_cfg{dim: C.uint32_t(dim), metric: C.int32_t(distance), quant: C.int32_t(quantization),
		device: C.int32_t(device), capacity_hint: C.uint64_t(capacityHint)}
	var h *C.coltt_store
	if err := call(func() C.int { return C.coltt_b200_store_create(&cfg, &h) }); err != nil {
		return nil, err
	}
	s := &Store{h: h, dim: int(dim)}
	runtime.SetFinalizer(s, func(s *Store) { s.Close() })
	return s, nil
}

func (s *Store) Close() {
	if s.h != nil {
		C.coltt_b200_store_destroy(s.h)
		s.h = nil
	}
}

// ChangedVertex: edge/none_vectorstore.go:66-103 (normalize + Lower happen on the GPU).
func (s *Store) ChangedVertex(id uint64, vector []float32) error {
	if len(vector) != s.dim {
		return errDim(s.dim, len(vector))
	}
	return call(func() C.int {
		return C.coltt_b200_store_upsert(s.h, (*C.uint64_t)(unsafe.Pointer(&id)), (*C.float)(unsafe.Pointer(&vector[0])), 1)
	})
}

// ChangedVertices is the batched form (bulk load).
func (s *Store) ChangedVertices(ids []uint64, vectors []float32) error {
	if len(ids) == 0 {
		return nil
	}
	if len(vectors) != len(ids)*s.dim {
		return errDim(len(ids)*s.dim, len(vectors))
	}
	return call(func() C.int {
		return C.coltt_b200_store_upsert(s.h, (*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&vectors[0])), C.size_t(len(ids)))
	})
}

// RemoveVertex after dropFilter was resolved to ids (none_vectorstore.go:105-127).
func (s *Store) RemoveVertex(ids []uint64) error {
	if len(ids) == 0 {
		return nil
	}
	return call(func() C.int {
		return C.coltt_b200_store_remove(s.h, (*C.uint64_t)(unsafe.Pointer(&ids[0])), C.size_t(len(ids)))
	})
}

// Hit is edge.SearchResultItem without the metadata map (the caller re-attaches it by Id).
type Hit struct {
	Id    uint64
	Score float32
}

// VertexSearch: edge/none_vectorstore.go:129-180.  highCpu has no meaning on the GPU.
func (s *Store) VertexSearch(target []float32, topK int, selectMode, mathMode int) ([]Hit, error) {
	if len(target) != s.dim {
		return nil, errDim(s.dim, len(target))
	}
	if topK <= 0 {
		return nil, errTopK
	}
	ids := make([]uint64, topK)
	scores := make([]float32, topK)
	var count C.int32_t
	if err := call(func() C.int {
		return C.coltt_b200_store_search(s.h, (*C.float)(unsafe.Pointer(&target[0])), 1, C.int(topK), C.int(selectMode), C.int(mathMode),
			(*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&scores[0])), &count)
	}); err != nil {
		return nil, err
	}
	out := make([]Hit, int(count))
	for i := range out {
		out[i] = Hit{ids[i], scores[i]}
	}
	return out, nil
}

// FilterableVertexSearch: edge/none_vectorstore.go:182-253, given inverted.SearchWithExpression's ids.
func (s *Store) FilterableVertexSearch(candidates []uint64, target []float32, topK int, selectMode int) ([]Hit, error) {
	if len(candidates) == 0 {
		return nil, nil
	}
	if len(target) != s.dim {
		return nil, errDim(s.dim, len(target))
	}
	if topK <= 0 {
		return nil, errTopK
	}
	ids := make([]uint64, topK)
	scores := make([]float32, topK)
	var count C.int32_t
	if err := call(func() C.int {
		return C.coltt_b200_store_search_subset(s.h, (*C.float)(unsafe.Pointer(&target[0])), 1,
			(*C.uint64_t)(unsafe.Pointer(&candidates[0])), C.size_t(len(candidates)), C.int(topK), C.int(selectMode),
			(*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&scores[0])), &count)
	}); err != nil {
		return nil, err
	}
	out := make([]Hit, int(count))
	for i := range out {
		out[i] = Hit{ids[i], scores[i]}
	}
	return out, nil
}

// BatchVertexSearch is the new surface a micro-batcher in Edge.Search (edge/edge.go:610-690) would call.
func (s *Store) BatchVertexSearch(targets []float32, nq, topK int, selectMode, mathMode int) ([]uint64, []float32, []int32, error) {
	if nq <= 0 {
		return nil, nil, nil, nil
	}
	if len(targets) != nq*s.dim {
		return nil, nil, nil, errDim(nq*s.dim, len(targets))
	}
	if topK <= 0 {
		return nil, nil, nil, errTopK
	}
	ids := make([]uint64, nq*topK)
	scores := make([]float32, nq*topK)
	counts := make([]int32, nq)
	err := call(func() C.int {
		return C.coltt_b200_store_search(s.h, (*C.float)(unsafe.Pointer(&targets[0])), C.size_t(nq), C.int(topK), C.int(selectMode), C.int(mathMode),
			(*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&scores[0])), (*C.int32_t)(unsafe.Pointer(&counts[0])))
	})
	return ids, scores, counts, err
}

// PinnedFloats is a []float32 over page-locked host memory (coltt_b200_host_alloc).  A query batch assembled in it is
// DMA'd to the GPU in place by BatchVertexSearch / Cluster.BatchVertexSearch; ordinary Go slices are first copied into the
// handle's own pinned staging buffer by the library.  C memory: not moved or scanned by the Go GC; Free it explicitly.
type PinnedFloats struct {
	Data []float32
	p    unsafe.Pointer
}

func HostAlloc(nFloats int) (*PinnedFloats, error) {
	if nFloats <= 0 {
		return nil, errors.New("colttb200: HostAlloc needs a positive size")
	}
	var p unsafe.Pointer
	if err := call(func() C.int { return C.coltt_b200_host_alloc(C.size_t(nFloats)*4, &p) }); err != nil {
		return nil, err
	}
	return &PinnedFloats{Data: unsafe.Slice((*float32)(p), nFloats), p: p}, nil
}

func (b *PinnedFloats) Free() {
	if b.p != nil {
		C.coltt_b200_host_free(b.p)
		b.p, b.Data = nil, nil
	}
}

// SaveVertex / LoadVertex: edge/none_vectorstore.go:308-516 (metaCount = 0; metadata is saved by the Go side).
func (s *Store) SaveVertex() ([]byte, error) {
	var n C.size_t
	if err := call(func() C.int { return C.coltt_b200_store_export(s.h, nil, &n) }); err != nil {
		return nil, err
	}
	buf := make([]byte, int(n)+1)
	if err := call(func() C.int { return C.coltt_b200_store_export(s.h, unsafe.Pointer(&buf[0]), &n) }); err != nil {
		return nil, err
	}
	return buf[:int(n)], nil
}

func (s *Store) LoadVertex(data []byte) error {
	if len(data) == 0 {
		return call(func() C.int { return C.coltt_b200_store_import(s.h, nil, 0) })
	}
	return call(func() C.int { return C.coltt_b200_store_import(s.h, unsafe.Pointer(&data[0]), C.size_t(len(data))) })
}

func (s *Store) LoadSize() (int64, error) {
	var n C.uint64_t
	err := call(func() C.int { return C.coltt_b200_store_size(s.h, &n) })
	return int64(n), err
}

// Hnsw wraps a device-resident vectorindex.Hnsw built from Hnsw.Commit(w, true).
type Hnsw struct {
	h   *C.coltt_hnsw
	dim int
}

func newHnsw(h *C.coltt_hnsw) (*Hnsw, error) {
	var d C.uint32_t
	if err := call(func() C.int { return C.coltt_b200_hnsw_dim(h, &d) }); err != nil {
		C.coltt_b200_hnsw_destroy(h)
		return nil, err
	}
	g := &Hnsw{h: h, dim: int(d)}
	runtime.SetFinalizer(g, func(g *Hnsw) { g.Close() })
	return g, nil
}

// LoadHnsw: core/vectorindex/hnsw_commit.go:164-278.
func LoadHnsw(commitBlob []byte, device int32) (*Hnsw, error) {
	if len(commitBlob) == 0 {
		return nil, errors.New("empty Hnsw.Commit blob")
	}
	var h *C.coltt_hnsw
	if err := call(func() C.int {
		return C.coltt_b200_hnsw_load(unsafe.Pointer(&commitBlob[0]), C.size_t(len(commitBlob)), C.int(device), &h)
	}); err != nil {
		return nil, err
	}
	return newHnsw(h)
}

func (g *Hnsw) Close() {
	if g.h != nil {
		C.coltt_b200_hnsw_destroy(g.h)
		g.h = nil
	}
}

// Search: core/vectorindex/hnsw.go:243-278 (ef <= 0 uses the ef stored in the blob).
func (g *Hnsw) Search(query []float32, k int, ef int) ([]Hit, error) {
	if len(query) != g.dim {
		return nil, errDim(g.dim, len(query))
	}
	if k <= 0 {
		return nil, errTopK
	}
	ids := make([]uint64, k)
	scores := make([]float32, k)
	var count C.int32_t
	if err := call(func() C.int {
		return C.coltt_b200_hnsw_search(g.h, (*C.float)(unsafe.Pointer(&query[0])), 1, C.int(k), C.int(ef),
			(*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&scores[0])), &count)
	}); err != nil {
		return nil, err
	}
	out := make([]Hit, int(count))
	for i := range out {
		out[i] = Hit{ids[i], scores[i]}
	}
	return out, nil
}

// BuildHnsw: bulk construction on the GPU in place of n x Hnsw.Insert (core/vectorindex/hnsw.go:104-167);
// levels may be nil (drawn like RandomLevel, hnsw.go:280-282).
func BuildHnsw(dim uint32, distance int32, m, ef, efConstruction int32, device int32, seed uint64, ids []uint64, vectors []float32, levels []int32) (*Hnsw, error) {
	cfg := C.coltt_hnsw_build_cfg{dim: C.uint32_t(dim), metric: C.int32_t(distance), m: C.int32_t(m), ef: C.int32_t(ef),
		ef_construction: C.int32_t(efConstruction), device: C.int32_t(device), seed: C.uint64_t(seed)}
	var lv *C.int32_t
	if len(levels) > 0 {
		lv = (*C.int32_t)(unsafe.Pointer(&levels[0]))
	}
	if len(ids) == 0 || len(vectors) != len(ids)*int(dim) || (len(levels) > 0 && len(levels) != len(ids)) {
		return nil, errDim(len(ids)*int(dim), len(vectors))
	}
	var h *C.coltt_hnsw
	if err := call(func() C.int {
		return C.coltt_b200_hnsw_build(&cfg, (*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&vectors[0])), lv, C.size_t(len(ids)), &h)
	}); err != nil {
		return nil, err
	}
	return newHnsw(h)
}

// Commit: Hnsw.Commit(w, true) (core/vectorindex/hnsw_commit.go:69-162).
func (g *Hnsw) Commit() ([]byte, error) {
	var n C.size_t
	if err := call(func() C.int { return C.coltt_b200_hnsw_commit(g.h, nil, &n) }); err != nil {
		return nil, err
	}
	buf := make([]byte, int(n)+1)
	if err := call(func() C.int { return C.coltt_b200_hnsw_commit(g.h, unsafe.Pointer(&buf[0]), &n) }); err != nil {
		return nil, err
	}
	return buf[:int(n)], nil
}

// MultiVertexSearch: experimental/multi_vector_vertex.go:85-137 over one Store per included vector field
// (request order); hits come back with descending Score like multi_priority_queue.go ToSlice().
func MultiVertexSearch(fields []*Store, queries [][]float32, ratios []int32, topK int) ([]Hit, error) {
	nf := len(fields)
	if nf == 0 || len(queries) != nf || len(ratios) != nf {
		return nil, errors.New("fields, queries and ratios must have the same non-zero length")
	}
	if topK <= 0 {
		return nil, errTopK
	}
	for j := range queries {
		if len(queries[j]) != fields[j].dim {
			return nil, errDim(fields[j].dim, len(queries[j]))
		}
	}
	// C arrays of handles / query pointers live in C memory: cgo forbids passing Go memory that holds Go pointers
	hs := (*[1 << 20]*C.coltt_store)(C.malloc(C.size_t(nf) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	qs := (*[1 << 20]*C.float)(C.malloc(C.size_t(nf) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(hs))
	defer C.free(unsafe.Pointer(qs))
	dim := len(queries[0])
	flat := (*C.float)(C.malloc(C.size_t(nf*dim) * 4))
	defer C.free(unsafe.Pointer(flat))
	fl := unsafe.Slice((*float32)(unsafe.Pointer(flat)), nf*dim)
	for j := 0; j < nf; j++ {
		hs[j] = fields[j].h
		copy(fl[j*dim:(j+1)*dim], queries[j])
		qs[j] = (*C.float)(unsafe.Pointer(&fl[j*dim]))
	}
	ids := make([]uint64, topK)
	scores := make([]float32, topK)
	var count C.int32_t
	if err := call(func() C.int {
		return C.coltt_b200_multi_search(&hs[0], &qs[0], (*C.int32_t)(unsafe.Pointer(&ratios[0])), C.int(nf), C.int(topK),
			(*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&scores[0])), &count)
	}); err != nil {
		return nil, err
	}
	out := make([]Hit, int(count))
	for i := range out {
		out[i] = Hit{ids[i], scores[i]}
	}
	return out, nil
}

// ---- sharded collections: one shard per GPU, one NCCL all-gather of per-shard top-k, merge ---------------------------
// Reference analogue: the 16 in-process map shards of a vectorspace re-merged after the scan
// (edge/none_vectorstore.go:152-178).  Rows go to GPU ShardVertex(id, 16) % nGPU (pkg/sharding/shard.go:34-41).

// Cluster is every GPU of this process: coltt_b200_init (ncclCommInitAll) + one Store per GPU.
type Cluster struct {
	comms  []*C.coltt_comm
	Shards []*Store
	dim    int
}

// NewCluster opens the communicators and one shard store per device.
func NewCluster(devices []int32, dim uint32, distance, quantization int32, capacityHintPerShard uint64) (*Cluster, error) {
	n := len(devices)
	if n == 0 {
		return nil, errors.New("no devices")
	}
	// C arrays live in C memory (cgo forbids handing C Go memory that holds pointers)
	devs := (*[64]C.int)(C.malloc(C.size_t(n) * 4))
	cms := (*[64]*C.coltt_comm)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(devs))
	defer C.free(unsafe.Pointer(cms))
	for i, d := range devices {
		devs[i] = C.int(d)
	}
	if err := call(func() C.int { return C.coltt_b200_init(&devs[0], C.int(n), &cms[0]) }); err != nil {
		return nil, err
	}
	c := &Cluster{dim: int(dim)}
	for i := 0; i < n; i++ {
		c.comms = append(c.comms, cms[i])
		s, err := NewStore(dim, distance, quantization, devices[i], capacityHintPerShard)
		if err != nil {
			return nil, err
		}
		c.Shards = append(c.Shards, s)
	}
	return c, nil
}

// ShardOf is the row -> GPU map: the reference's shard identity folded onto the GPUs.
func (c *Cluster) ShardOf(id uint64) int {
	h := uint64(14695981039346656037)
	for i := 0; i < 8; i++ {
		h ^= (id >> (8 * uint(i))) & 0xff
		h *= 1099511628211
	}
	return int((h % 16) % uint64(len(c.Shards)))
}

// BatchVertexSearch answers nq queries over all shards: coltt_b200_sharded_search_all runs one library thread per
// rank (local search -> ncclAllGather -> merge) and returns rank 0's merged answer.
func (c *Cluster) BatchVertexSearch(targets []float32, nq, topK int, selectMode, mathMode int) ([]uint64, []float32, []int32, error) {
	if nq <= 0 {
		return nil, nil, nil, nil
	}
	if len(targets) != nq*c.dim {
		return nil, nil, nil, errDim(nq*c.dim, len(targets))
	}
	if topK <= 0 {
		return nil, nil, nil, errTopK
	}
	n := len(c.comms)
	cms := (*[64]*C.coltt_comm)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	shs := (*[64]*C.coltt_store)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(cms))
	defer C.free(unsafe.Pointer(shs))
	for i := 0; i < n; i++ {
		cms[i] = c.comms[i]
		shs[i] = c.Shards[i].h
	}
	ids := make([]uint64, nq*topK)
	scores := make([]float32, nq*topK)
	counts := make([]int32, nq)
	err := call(func() C.int {
		return C.coltt_b200_sharded_search_all(&cms[0], &shs[0], C.int(n), (*C.float)(unsafe.Pointer(&targets[0])), C.size_t(nq), C.int(topK),
			C.int(selectMode), C.int(mathMode), (*C.uint64_t)(unsafe.Pointer(&ids[0])), (*C.float)(unsafe.Pointer(&scores[0])),
			(*C.int32_t)(unsafe.Pointer(&counts[0])))
	})
	return ids, scores, counts, err
}

// Close destroys the shard stores and every communicator of the process.
func (c *Cluster) Close() {
	for _, s := range c.Shards {
		s.Close()
	}
	C.coltt_b200_shutdown()
	c.comms = nil
}
