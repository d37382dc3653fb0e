// UNCOMPILED (no Go toolchain here).  Same factory symbol the core looks up with plugin.Open
// (reference: gpuschedulerplugin/plugin/gpuscheduler.go:8).
package main

import (
	"github.com/Microsoft/KubeDevice-API/pkg/devicescheduler"
	"github.com/Microsoft/KubeGPU/go/kgpuscheduler"
)

func CreateDeviceSchedulerPlugin() (devicescheduler.DeviceScheduler, error) {
	return kgpuscheduler.New([]int32{0})
}
