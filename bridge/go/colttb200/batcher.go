// batcher.go — micro-batcher for the single-vector Search RPC (edge/edge.go:610-690): every RPC goroutine
// calls Batcher.VertexSearch and blocks; one flusher goroutine coalesces up to MaxBatch waiting queries of equal
// topK (or whatever has arrived MaxWait after the first one) into ONE Store.BatchVertexSearch.  Go twin of
// coltt_b200/batcher.py; written but not compiled here (no Go toolchain in the build image).
package colttb200

import "time"

type batchReq struct {
	q    []float32
	topK int
	out  chan batchRes
}
type batchRes struct {
	hits []Hit
	err  error
}

type Batcher struct {
	store            *Store
	dim              int
	MaxBatch         int
	MaxWait          time.Duration
	selectMode, math int
	in               chan batchReq
}

func NewBatcher(s *Store, dim, maxBatch int, maxWait time.Duration, selectMode, mathMode int) *Batcher {
	b := &Batcher{store: s, dim: dim, MaxBatch: maxBatch, MaxWait: maxWait, selectMode: selectMode, math: mathMode, in: make(chan batchReq, 4*maxBatch)}
	go b.run()
	return b
}

// VertexSearch: what Edge.Search calls instead of Vectorstore.VertexSearch (edge/edge.go:653).
func (b *Batcher) VertexSearch(target []float32, topK int) ([]Hit, error) {
	// validated here, before the request can reach the shared batch: one wrong-length query would shift every later
	// query of the flattened batch, and topK <= 0 would panic the single flusher goroutine
	if len(target) != b.dim {
		return nil, errDim(b.dim, len(target))
	}
	if topK <= 0 {
		return nil, errTopK
	}
	r := batchReq{q: target, topK: topK, out: make(chan batchRes, 1)}
	b.in <- r
	res := <-r.out
	return res.hits, res.err
}

func (b *Batcher) run() {
	var held []batchReq // requests whose topK differs from the batch being formed
	pinned, _ := HostAlloc(b.MaxBatch * b.dim) // nil on failure: ordinary slices still work, through the library's staging copy
	for {
		var first batchReq
		if len(held) > 0 {
			first, held = held[0], held[1:]
		} else {
			first = <-b.in
		}
		batch := []batchReq{first}
		timer := time.NewTimer(b.MaxWait)
	fill:
		for len(batch) < b.MaxBatch {
			select {
			case r := <-b.in:
				if r.topK == first.topK {
					batch = append(batch, r)
				} else {
					held = append(held, r)
				}
			case <-timer.C:
				break fill
			}
		}
		timer.Stop()
		// the batch is assembled in page-locked memory (one buffer for the life of the batcher): the library DMAs it in place
		flat := make([]float32, 0)
		if pinned != nil {
			flat = pinned.Data[:0]
		}
		for _, r := range batch {
			flat = append(flat, r.q...)
		}
		ids, scores, counts, err := b.store.BatchVertexSearch(flat, len(batch), first.topK, b.selectMode, b.math)
		for j, r := range batch {
			if err != nil {
				r.out <- batchRes{nil, err}
				continue
			}
			hits := make([]Hit, int(counts[j]))
			for i := range hits {
				hits[i] = Hit{ids[j*first.topK+i], scores[j*first.topK+i]}
			}
			r.out <- batchRes{hits, nil}
		}
	}
}
