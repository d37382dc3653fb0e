// Package kgpuscheduler: cgo binding of include/kgpu.h.  UNCOMPILED (no Go toolchain here).
package kgpuscheduler

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../kubegpu_b200/lib -lkgpu
#include <stdlib.h>
#include "kgpu.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// NoFit mirrors KGPU_NO_FIT.
const NoFit = ^uint64(0)

type handle struct{ h *C.kgpu_t }

func create(devs []int32) (*handle, error) {
	var h *C.kgpu_t
	if rc := C.kgpu_create((*C.int)(unsafe.Pointer(&devs[0])), C.int(len(devs)), &h); rc != C.KGPU_OK {
		return nil, fmt.Errorf("kgpu_create: %s", C.GoString(C.kgpu_last_error(nil)))
	}
	return &handle{h}, nil
}

func (k *handle) destroy() { C.kgpu_destroy(k.h) }

func (k *handle) err(what string) error {
	return fmt.Errorf("%s: %s", what, C.GoString(C.kgpu_last_error(k.h)))
}

// uploadNodes: topo is N*64 int32, free is N int32.  The slices are only borrowed for the
// duration of the call (cgo pointer rule); libkgpu copies them to the device before returning.
func (k *handle) uploadNodes(topo, free []int32, base int64) error {
	if len(free) == 0 {
		if rc := C.kgpu_upload_nodes(k.h, nil, nil, 0, C.int64_t(base)); rc != C.KGPU_OK {
			return k.err("kgpu_upload_nodes")
		}
		return nil
	}
	if rc := C.kgpu_upload_nodes(k.h, (*C.int32_t)(unsafe.Pointer(&topo[0])),
		(*C.int32_t)(unsafe.Pointer(&free[0])), C.int64_t(len(free)), C.int64_t(base)); rc != C.KGPU_OK {
		return k.err("kgpu_upload_nodes")
	}
	return nil
}

func (k *handle) updateNode(idx int64, topo *[64]int32, free int32) error {
	if rc := C.kgpu_update_node(k.h, C.int64_t(idx), (*C.int32_t)(unsafe.Pointer(&topo[0])), C.int32_t(free)); rc != C.KGPU_OK {
		return k.err("kgpu_update_node")
	}
	return nil
}

func (k *handle) setFreeMask(idx int64, free int32) error {
	if rc := C.kgpu_set_free_mask(k.h, C.int64_t(idx), C.int32_t(free)); rc != C.KGPU_OK {
		return k.err("kgpu_set_free_mask")
	}
	return nil
}

// setFreeMasks: a cycle's Take/Return in ONE call (one H2D copy + one kernel on the device).
func (k *handle) setFreeMasks(idx []int64, free []int32) error {
	if len(idx) == 0 {
		return nil
	}
	if rc := C.kgpu_set_free_masks(k.h, (*C.int64_t)(unsafe.Pointer(&idx[0])), (*C.int32_t)(unsafe.Pointer(&free[0])),
		C.int64_t(len(idx))); rc != C.KGPU_OK {
		return k.err("kgpu_set_free_masks")
	}
	return nil
}

// fitLookup: (cost<<8 | mask) of one node for one GPU count from the handle's host-side (node, k)
// table -- no launch, no copy: this is what PodFitsDevice calls once per (node, pod) pair.
func (k *handle) fitLookup(node int64, gpus int32) (uint32, error) {
	var out C.uint32_t
	if rc := C.kgpu_fit_lookup(k.h, C.int64_t(node), C.int32_t(gpus), &out); rc != C.KGPU_OK {
		return 0, k.err("kgpu_fit_lookup")
	}
	return uint32(out), nil
}

// scoreBatch: pods is P*4 int32 {k, pod_id, flags, reserved}; returns P keys
// (cost<<40 | node<<8 | mask) or NoFit.
func (k *handle) scoreBatch(pods []int32) ([]uint64, error) {
	keys := make([]uint64, len(pods)/4)
	if len(keys) == 0 {
		return keys, nil
	}
	if rc := C.kgpu_score_batch(k.h, (*C.int32_t)(unsafe.Pointer(&pods[0])), C.int64_t(len(keys)),
		(*C.uint64_t)(unsafe.Pointer(&keys[0]))); rc != C.KGPU_OK {
		return nil, k.err("kgpu_score_batch")
	}
	return keys, nil
}

// scorePair: (cost<<8 | mask) of one node for one GPU count, or 0xFFFFFFFF.
func (k *handle) scorePair(node int64, gpus int32) (uint32, error) {
	var out C.uint32_t
	n, kk := C.int64_t(node), C.int32_t(gpus)
	if rc := C.kgpu_score_pairs(k.h, &n, &kk, nil, 1, &out); rc != C.KGPU_OK {
		return 0, k.err("kgpu_score_pairs")
	}
	return uint32(out), nil
}

// placeBatch: stateful sequential placement (pods in order, the device-side free masks are
// updated).  dryRun: conflict-free proposals on a scratch copy of the masks, nothing is taken.
// Same buffers as scoreBatch.
func (k *handle) placeBatch(pods []int32, dryRun bool) ([]uint64, error) {
	keys := make([]uint64, len(pods)/4)
	if len(keys) == 0 {
		return keys, nil
	}
	flags := C.int(0)
	if dryRun {
		flags = C.KGPU_PLACE_DRY_RUN
	}
	if rc := C.kgpu_place_batch_ex(k.h, (*C.int32_t)(unsafe.Pointer(&pods[0])), C.int64_t(len(keys)),
		(*C.uint64_t)(unsafe.Pointer(&keys[0])), flags); rc != C.KGPU_OK {
		return nil, k.err("kgpu_place_batch_ex")
	}
	return keys, nil
}

func (k *handle) freeMasks(n int64) ([]int32, error) {
	out := make([]int32, n)
	if n == 0 {
		return out, nil
	}
	if rc := C.kgpu_get_free_masks(k.h, (*C.int32_t)(unsafe.Pointer(&out[0])), C.int64_t(n)); rc != C.KGPU_OK {
		return nil, k.err("kgpu_get_free_masks")
	}
	return out, nil
}
