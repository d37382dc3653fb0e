// UNCOMPILED (no Go toolchain here).  DeviceScheduler over libkgpu with the reference's
// method set (gpuschedulerplugin/gpu_scheduler.go:21-71).  The tree cache / DevRequests rewrite
// is delegated to the reference's own package so the core's group allocator keeps working;
// the score and the placement come from the GPU.
package kgpuscheduler

import (
	"fmt"
	"regexp"
	"sort"
	"sync"

	"github.com/Microsoft/KubeDevice-API/pkg/devicescheduler"
	types "github.com/Microsoft/KubeDevice-API/pkg/types"
	gtype "github.com/Microsoft/KubeGPU/gpuplugintypes"
	ref "github.com/Microsoft/KubeGPU/gpuschedulerplugin"
)

var gpuKey = regexp.MustCompile(string(types.DeviceGroupPrefix) + `/(gpugrp1/(.*?)/gpugrp0/(.*?)/gpu/(.*?))/cards`)

type nodeRec struct {
	index   int64
	names   []string // slot -> "gpugrp1/a/gpugrp0/b/gpu/<id>"
	topo    [64]int32
	present uint32
	used    uint32
	removed bool
}

type placement struct {
	node string
	mask uint32
	cost uint32
}

// NvidiaGPUScheduler implements devicescheduler.DeviceScheduler.
type NvidiaGPUScheduler struct {
	mu     sync.Mutex
	h      *handle
	nodes  map[string]*nodeRec
	byIdx  []string
	dirty  bool
	placed map[string]placement // pod name -> last ScoreBatch result
}

// New creates the scheduler on the given CUDA devices (one handle shards over all of them).
func New(devs []int32) (*NvidiaGPUScheduler, error) {
	h, err := create(devs)
	if err != nil {
		return nil, err
	}
	return &NvidiaGPUScheduler{h: h, nodes: map[string]*nodeRec{}, placed: map[string]placement{}}, nil
}

func level(a, b []string) int32 { // same gpugrp0 -> 5, same gpugrp1 -> 3, else 1
	if a[2] == b[2] && a[3] == b[3] {
		return 5
	}
	if a[2] == b[2] {
		return 3
	}
	return 1
}

func (ns *NvidiaGPUScheduler) AddNode(nodeName string, nodeInfo *types.NodeInfo) {
	(&ref.NvidiaGPUScheduler{}).AddNode(nodeName, nodeInfo) // translation + tree cache, gpu_scheduler.go:21-28
	ns.mu.Lock()
	defer ns.mu.Unlock()
	keys := make([]string, 0, len(nodeInfo.Allocatable))
	for k := range nodeInfo.Allocatable {
		keys = append(keys, string(k))
	}
	sort.Strings(keys)
	var slots [][]string
	for _, k := range keys {
		if m := gpuKey.FindStringSubmatch(k); len(m) == 5 && len(slots) < 8 {
			slots = append(slots, m)
		}
	}
	rec, ok := ns.nodes[nodeName]
	if !ok {
		rec = &nodeRec{index: int64(len(ns.byIdx))}
		ns.nodes[nodeName] = rec
		ns.byIdx = append(ns.byIdx, nodeName)
	}
	rec.removed, rec.names, rec.topo = false, nil, [64]int32{}
	for i, a := range slots {
		rec.names = append(rec.names, a[1])
		for j, b := range slots {
			if i != j {
				rec.topo[i*8+j] = level(a, b)
			}
		}
	}
	rec.present = uint32(1)<<uint(len(slots)) - 1
	ns.dirty = true
}

func (ns *NvidiaGPUScheduler) RemoveNode(nodeName string) {
	(&ref.NvidiaGPUScheduler{}).RemoveNode(nodeName)
	ns.mu.Lock()
	defer ns.mu.Unlock()
	if rec, ok := ns.nodes[nodeName]; ok {
		rec.removed, ns.dirty = true, true
	}
}

func (ns *NvidiaGPUScheduler) flush() error {
	if !ns.dirty {
		return nil
	}
	topo := make([]int32, 64*len(ns.byIdx))
	free := make([]int32, len(ns.byIdx))
	for i, name := range ns.byIdx {
		rec := ns.nodes[name]
		copy(topo[64*i:], rec.topo[:])
		if !rec.removed {
			free[i] = int32(rec.present &^ rec.used)
		}
	}
	if err := ns.h.uploadNodes(topo, free, 0); err != nil {
		return err
	}
	ns.dirty = false
	return nil
}

func podGPUs(podInfo *types.PodInfo) int32 { // gpu.go:295-303
	n := int64(0)
	for _, c := range podInfo.RunningContainers {
		n += c.Requests[gtype.ResourceGPU]
	}
	for _, c := range podInfo.InitContainers {
		if c.Requests[gtype.ResourceGPU] > n {
			n = c.Requests[gtype.ResourceGPU]
		}
	}
	if n > 8 {
		n = 9
	}
	return int32(n)
}

// ScoreBatch is the batched side door: one kernel launch for a whole scheduling cycle.
func (ns *NvidiaGPUScheduler) ScoreBatch(pods []*types.PodInfo) ([]uint64, error) {
	ns.mu.Lock()
	defer ns.mu.Unlock()
	if err := ns.flush(); err != nil {
		return nil, err
	}
	req := make([]int32, 4*len(pods))
	for i, p := range pods {
		req[4*i], req[4*i+1] = podGPUs(p), int32(i)
	}
	keys, err := ns.h.scoreBatch(req)
	if err != nil {
		return nil, err
	}
	for i, p := range pods {
		if keys[i] != NoFit {
			ns.placed[p.Name] = placement{ns.byIdx[(keys[i]>>8)&0xFFFFFFFF], uint32(keys[i] & 0xFF), uint32(keys[i] >> 40)}
		}
	}
	return keys, nil
}

func (ns *NvidiaGPUScheduler) PodFitsDevice(nodeInfo *types.NodeInfo, podInfo *types.PodInfo, fillAllocateFrom bool) (bool, []devicescheduler.PredicateFailureReason, float64) {
	fits, reasons, _ := (&ref.NvidiaGPUScheduler{}).PodFitsDevice(nodeInfo, podInfo, fillAllocateFrom)
	if !fits {
		return false, reasons, 0.0
	}
	ns.mu.Lock()
	defer ns.mu.Unlock()
	rec, ok := ns.nodes[nodeInfo.Name]
	if !ok || rec.removed || ns.flush() != nil {
		return true, nil, 0.0
	}
	nk, err := ns.h.scorePair(rec.index, podGPUs(podInfo))
	if err != nil {
		return true, nil, 0.0
	}
	if nk == 0xFFFFFFFF {
		return false, nil, 0.0
	}
	return true, nil, 1.0 / (1.0 + float64(nk>>8))
}

func (ns *NvidiaGPUScheduler) PodAllocate(nodeInfo *types.NodeInfo, podInfo *types.PodInfo) error {
	if err := (&ref.NvidiaGPUScheduler{}).PodAllocate(nodeInfo, podInfo); err != nil {
		return err
	}
	ns.mu.Lock()
	defer ns.mu.Unlock()
	pl, ok := ns.placed[podInfo.Name]
	if !ok {
		return nil
	}
	rec := ns.nodes[pl.node]
	var slots []int
	for i := 0; i < 8; i++ {
		if pl.mask>>uint(i)&1 == 1 {
			slots = append(slots, i)
		}
	}
	next := 0
	names := make([]string, 0, len(podInfo.RunningContainers))
	for n := range podInfo.RunningContainers {
		names = append(names, n)
	}
	sort.Strings(names)
	for _, n := range names {
		cont := podInfo.RunningContainers[n]
		cont.AllocateFrom = types.ResourceLocation{}
		reqs := make([]string, 0, len(cont.DevRequests))
		for r := range cont.DevRequests {
			reqs = append(reqs, string(r))
		}
		sort.Strings(reqs)
		for _, r := range reqs {
			if next < len(slots) && regexp.MustCompile(`/gpu/.*/cards$`).MatchString(r) {
				cont.AllocateFrom[types.ResourceName(r)] = types.ResourceName(fmt.Sprintf("%s/%s/cards", types.DeviceGroupPrefix, rec.names[slots[next]]))
				next++
			}
		}
		podInfo.RunningContainers[n] = cont
	}
	return nil
}

func (ns *NvidiaGPUScheduler) take(podInfo *types.PodInfo, release bool) error {
	ns.mu.Lock()
	defer ns.mu.Unlock()
	pl, ok := ns.placed[podInfo.Name]
	if !ok {
		return nil
	}
	rec := ns.nodes[pl.node]
	if release {
		rec.used &^= pl.mask
		delete(ns.placed, podInfo.Name)
	} else {
		if rec.used&pl.mask != 0 {
			return fmt.Errorf("TakePodResources: GPUs already in use on %s", pl.node)
		}
		rec.used |= pl.mask
	}
	if ns.dirty {
		return nil
	}
	return ns.h.setFreeMask(rec.index, int32(rec.present&^rec.used))
}

func (ns *NvidiaGPUScheduler) TakePodResources(nodeInfo *types.NodeInfo, podInfo *types.PodInfo) error {
	return ns.take(podInfo, false)
}

func (ns *NvidiaGPUScheduler) ReturnPodResources(nodeInfo *types.NodeInfo, podInfo *types.PodInfo) error {
	return ns.take(podInfo, true)
}

func (ns *NvidiaGPUScheduler) GetName() string { return "nvidiagpu" }

func (ns *NvidiaGPUScheduler) UsingGroupScheduler() bool { return true }
