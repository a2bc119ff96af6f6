// Package cuda implements polaris's tracer.Tracer interface (tracer/tracer.go:80-111) over
// libpolaris_cuda.so, the B200 (sm_100a) path-tracing backend.  It is the cgo counterpart of
// tracer/opencl: renderer/default.go, renderer/opengl.go, the block schedulers of
// tracer/scheduler.go and the `polaris render` commands drive it unchanged once
// renderer.initTracers constructs these tracers instead of opencl ones (see INTEGRATION.md).
//
// NOTE: written against include/polaris_cuda.h without a Go toolchain in the build image; the
// same call sequence is exercised through ctypes by polaris_b200/tracer.py and the test suite.
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -lpolaris_cuda
#include <stdlib.h>
#include "polaris_cuda.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"image"
	"image/png"
	"math/rand"
	"os"
	"runtime"
	"sync"
	"time"
	"unsafe"

	"github.com/achilleasa/polaris/asset/scene"
	"github.com/achilleasa/polaris/tracer"
)

// Errors mirror tracer/opencl/errors.go so callers can keep comparing against sentinel values.
var (
	ErrNoDevice              = errors.New("cuda tracer: no usable CUDA device")
	ErrAllocatingBuffer      = errors.New("cuda tracer: could not allocate device buffer")
	ErrCopyingDataToHost     = errors.New("cuda tracer: could not copy device data to host buffer")
	ErrCopyingDataToDevice   = errors.New("cuda tracer: could not copy host data to device buffer")
	ErrKernelExecutionFailed = errors.New("cuda tracer: kernel execution failed")
	ErrUnsupportedChangeType = errors.New("cuda tracer: unsupported change type")
	ErrInvalidChangeData     = errors.New("cuda tracer: invalid data type for change")
	ErrNoSceneData           = errors.New("cuda tracer: no scene data uploaded")
)

// DeviceInfo is what `polaris list-devices` prints per device (cmd/list_devices.go:13-38).
type DeviceInfo struct {
	Ordinal  int
	Name     string
	SMs      uint32
	ClockMHz uint32
	Speed    uint32 // SMs*MHz/1000, the unit of tracer/opencl/device/device.go:209-222
}

// Devices enumerates the CUDA devices the library can drive.
func Devices() []DeviceInfo {
	n := int(C.pc_device_count())
	out := make([]DeviceInfo, 0, n)
	for i := 0; i < n; i++ {
		var name [256]C.char
		var sm, mhz, speed C.uint32_t
		if C.pc_device_info(C.int(i), &name[0], 256, &sm, &mhz, &speed) != 0 {
			continue
		}
		out = append(out, DeviceInfo{i, C.GoString(&name[0]), uint32(sm), uint32(mhz), uint32(speed)})
	}
	return out
}

// PostProcessStage is the cuda backend's counterpart of opencl.PipelineStage for the post-process list
// (tracer/opencl/pipeline.go:33-53): run by SyncFramebuffer after the tonemap, in order.
type PostProcessStage func(tr *Tracer, blockReq *tracer.BlockRequest) (time.Duration, error)

// SaveFrameBuffer writes the RGBA8 frame as a PNG, like opencl.SaveFrameBuffer (pipeline.go:216-236).
func SaveFrameBuffer(imgFile string) PostProcessStage {
	return func(tr *Tracer, blockReq *tracer.BlockRequest) (time.Duration, error) {
		start := time.Now()
		f, err := os.Create(imgFile)
		if err != nil {
			return 0, err
		}
		defer f.Close()
		im := &image.RGBA{Pix: tr.frameBuffer, Stride: int(blockReq.FrameW) * 4,
			Rect: image.Rect(0, 0, int(blockReq.FrameW), int(blockReq.FrameH))}
		if err := png.Encode(f, im); err != nil {
			return 0, err
		}
		return time.Since(start), nil
	}
}

// Tracer is one CUDA device behind the tracer.Tracer interface.
type Tracer struct {
	sync.Mutex
	// Stages run by SyncFramebuffer after the tonemap (renderer/cuda_backend.go fills it from Options).
	PostProcess []PostProcessStage
	id      string
	ordinal int
	handle  *C.pc_tracer
	stats   *tracer.Stats

	// queued state changes, applied by the next Trace (tracer/opencl/tracer.go:150-158,198)
	changeBuffer map[tracer.ChangeType]interface{}
	hasScene     bool

	// Seeds, when non-nil, supplies the host seed list (1 camera seed + NumBounces shade seeds per
	// sample, in the order tracer.go:222 / pipeline.go:146 draw them); nil draws from math/rand like
	// the reference.
	Seeds func(n int) []uint32

	frameBuffer []byte
}

// NewTracer mirrors opencl.NewTracer (tracer/opencl/tracer.go:58-73).
func NewTracer(id string, ordinal int) (*Tracer, error) {
	return &Tracer{
		id:           id,
		ordinal:      ordinal,
		stats:        &tracer.Stats{},
		changeBuffer: make(map[tracer.ChangeType]interface{}),
	}, nil
}

func (tr *Tracer) lastError(code C.int) error {
	msg := "unknown error"
	if m := C.pc_last_error(tr.handle); m != nil {
		msg = C.GoString(m)
	}
	var base error
	switch code {
	case C.PC_ERR_NO_DEVICE:
		base = ErrNoDevice
	case C.PC_ERR_ALLOC:
		base = ErrAllocatingBuffer
	case C.PC_ERR_COPY_TO_DEVICE:
		base = ErrCopyingDataToDevice
	case C.PC_ERR_COPY_TO_HOST:
		base = ErrCopyingDataToHost
	case C.PC_ERR_KERNEL:
		base = ErrKernelExecutionFailed
	case C.PC_ERR_NO_SCENE_DATA:
		return ErrNoSceneData
	default:
		return fmt.Errorf("cuda tracer: %s (code %d)", msg, int(code))
	}
	return fmt.Errorf("%w: %s", base, msg)
}

// Id, Flags, Speed: tracer/opencl/tracer.go:75-92.
func (tr *Tracer) Id() string         { return tr.id }
func (tr *Tracer) Flags() tracer.Flag { return tracer.Local }
func (tr *Tracer) Speed() uint32 {
	if tr.handle == nil {
		for _, d := range Devices() {
			if d.Ordinal == tr.ordinal {
				return d.Speed
			}
		}
		return 0
	}
	return uint32(C.pc_speed(tr.handle))
}

// Init: tracer/opencl/tracer.go:95-117.
func (tr *Tracer) Init() error {
	tr.Lock()
	defer tr.Unlock()
	if tr.handle != nil {
		return nil
	}
	cid := C.CString(tr.id)
	defer C.free(unsafe.Pointer(cid))
	if rc := C.pc_create(C.int(tr.ordinal), cid, &tr.handle); rc != 0 {
		msg := ""
		if m := C.pc_last_error(nil); m != nil {
			msg = C.GoString(m)
		}
		return fmt.Errorf("%w: %s", ErrNoDevice, msg)
	}
	return nil
}

// Close: tracer/opencl/tracer.go:120-142 (idempotent).
func (tr *Tracer) Close() {
	tr.Lock()
	defer tr.Unlock()
	if tr.handle != nil {
		C.pc_destroy(tr.handle)
		tr.handle = nil
	}
	tr.hasScene = false
}

func (tr *Tracer) Stats() *tracer.Stats { return tr.stats }

// UpdateState: tracer/opencl/tracer.go:150-158.
func (tr *Tracer) UpdateState(mode tracer.UpdateMode, changeType tracer.ChangeType, data interface{}) (time.Duration, error) {
	tr.Lock()
	defer tr.Unlock()
	tr.changeBuffer[changeType] = data
	if mode == tracer.Synchronous {
		return tr.commitChanges()
	}
	return 0, nil
}

// commitChanges: tracer/opencl/tracer.go:161-191.  The library copies every buffer during the
// call and never retains a Go pointer; see uploadScene for how the scene view satisfies cgo's pointer-passing rules.
func (tr *Tracer) commitChanges() (time.Duration, error) {
	if len(tr.changeBuffer) == 0 {
		return 0, nil
	}
	start := time.Now()
	for changeType, data := range tr.changeBuffer {
		switch changeType {
		case tracer.FrameDimensions:
			dims, ok := data.([2]uint32)
			if !ok {
				return time.Since(start), ErrInvalidChangeData
			}
			if rc := C.pc_resize(tr.handle, C.uint32_t(dims[0]), C.uint32_t(dims[1])); rc != 0 {
				return time.Since(start), tr.lastError(rc)
			}
			tr.frameBuffer = make([]byte, int(dims[0])*int(dims[1])*4)
		case tracer.SceneData:
			sc, ok := data.(*scene.Scene)
			if !ok {
				return time.Since(start), ErrInvalidChangeData
			}
			if err := tr.uploadScene(sc); err != nil {
				return time.Since(start), err
			}
		case tracer.CameraData:
			cam, ok := data.(*scene.Camera)
			if !ok {
				return time.Since(start), ErrInvalidChangeData
			}
			eye := [3]C.float{C.float(cam.Position[0]), C.float(cam.Position[1]), C.float(cam.Position[2])}
			var fr [16]C.float
			for k := 0; k < 4; k++ {
				for c := 0; c < 4; c++ {
					fr[4*k+c] = C.float(cam.Frustrum[k][c])
				}
			}
			if rc := C.pc_set_camera(tr.handle, &eye[0], &fr[0]); rc != 0 {
				return time.Since(start), tr.lastError(rc)
			}
		default:
			return time.Since(start), fmt.Errorf("%w %d", ErrUnsupportedChangeType, changeType)
		}
	}
	tr.changeBuffer = make(map[tracer.ChangeType]interface{})
	tr.stats.UpdateTime = time.Since(start)
	return tr.stats.UpdateTime, nil
}

// uploadScene: bufferSet.UploadSceneData (tracer/opencl/buffers.go:177-201); struct sizes are the
// ones asset/scene/optimized_scene.go documents (32/80/64/80/16 bytes).
//
// cgo pointer passing: pc_scene_view is a struct OF pointers into Go slices.  Handing C a Go-allocated struct
// that holds unpinned Go pointers panics under the default cgocheck ("cgo argument has Go pointer to unpinned
// Go pointer").  Every backing array is therefore pinned with runtime.Pinner (Go >= 1.21) for the duration of
// the call, and the view itself lives in C memory; the library copies during the call and retains nothing, so
// everything is unpinned / freed on return.
func (tr *Tracer) uploadScene(sc *scene.Scene) error {
	var pinner runtime.Pinner
	defer pinner.Unpin()
	v := (*C.pc_scene_view)(C.calloc(1, C.size_t(unsafe.Sizeof(C.pc_scene_view{}))))
	if v == nil {
		return ErrAllocatingBuffer
	}
	defer C.free(unsafe.Pointer(v))
	// pin(&slice[0]) and return (pointer, byte length); empty slices stay (nil, 0)
	view := func(first unsafe.Pointer, pin func(), n int, elem uintptr) (unsafe.Pointer, C.uint64_t) {
		if n == 0 {
			return nil, 0
		}
		pin()
		return first, C.uint64_t(uintptr(n) * elem)
	}
	if n := len(sc.BvhNodeList); n > 0 {
		p := &sc.BvhNodeList[0]
		v.bvh_nodes, v.bvh_nodes_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, unsafe.Sizeof(*p))
	}
	if n := len(sc.MeshInstanceList); n > 0 {
		p := &sc.MeshInstanceList[0]
		v.mesh_instances, v.mesh_instances_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, unsafe.Sizeof(*p))
	}
	if n := len(sc.MaterialNodeList); n > 0 {
		p := &sc.MaterialNodeList[0]
		v.material_nodes, v.material_nodes_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, unsafe.Sizeof(*p))
	}
	if n := len(sc.TextureData); n > 0 {
		p := &sc.TextureData[0]
		v.texture_data, v.texture_data_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, 1)
	}
	if n := len(sc.TextureMetadata); n > 0 {
		p := &sc.TextureMetadata[0]
		v.texture_metadata, v.texture_metadata_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, unsafe.Sizeof(*p))
	}
	if n := len(sc.VertexList); n > 0 {
		p := &sc.VertexList[0]
		v.vertices, v.vertices_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, unsafe.Sizeof(*p))
	}
	if n := len(sc.NormalList); n > 0 {
		p := &sc.NormalList[0]
		v.normals, v.normals_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, unsafe.Sizeof(*p))
	}
	if n := len(sc.UvList); n > 0 {
		p := &sc.UvList[0]
		v.uvs, v.uvs_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, unsafe.Sizeof(*p))
	}
	if n := len(sc.MaterialIndex); n > 0 {
		p := &sc.MaterialIndex[0]
		v.material_indices, v.material_indices_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, 4)
	}
	if n := len(sc.EmissivePrimitives); n > 0 {
		p := &sc.EmissivePrimitives[0]
		v.emissives, v.emissives_bytes = view(unsafe.Pointer(p), func() { pinner.Pin(p) }, n, unsafe.Sizeof(*p))
	}
	v.scene_diffuse_mat_index = C.int32_t(sc.SceneDiffuseMatIndex)
	v.scene_emissive_mat_index = C.int32_t(sc.SceneEmissiveMatIndex)
	if rc := C.pc_upload_scene(tr.handle, v); rc != 0 {
		return tr.lastError(rc)
	}
	tr.hasScene = true
	return nil
}

func toC(r *tracer.BlockRequest) C.pc_block_request {
	return C.pc_block_request{
		frame_w: C.uint32_t(r.FrameW), frame_h: C.uint32_t(r.FrameH),
		block_x: C.uint32_t(r.BlockX), block_y: C.uint32_t(r.BlockY),
		block_w: C.uint32_t(r.BlockW), block_h: C.uint32_t(r.BlockH),
		samples_per_pixel: C.uint32_t(r.SamplesPerPixel), num_bounces: C.uint32_t(r.NumBounces),
		min_bounces_for_rr: C.uint32_t(r.MinBouncesForRR), exposure: C.float(r.Exposure),
		seed: C.uint32_t(r.Seed), accumulated_samples: C.uint32_t(r.AccumulatedSamples),
	}
}

// Trace: tracer/opencl/tracer.go:194-247.  Seed and AccumulatedSamples are updated in place like
// the reference does.
func (tr *Tracer) Trace(blockReq *tracer.BlockRequest) (time.Duration, error) {
	tr.Lock()
	defer tr.Unlock()
	start := time.Now()
	if _, err := tr.commitChanges(); err != nil {
		return time.Since(start), err
	}
	if !tr.hasScene {
		return time.Since(start), ErrNoSceneData
	}
	n := int(blockReq.SamplesPerPixel) * (1 + int(blockReq.NumBounces))
	var seeds []uint32
	if tr.Seeds != nil {
		seeds = tr.Seeds(n)
	} else {
		seeds = make([]uint32, n)
		for i := range seeds {
			seeds[i] = rand.Uint32() // tracer.go:222, pipeline.go:146
		}
	}
	req := toC(blockReq)
	var st C.pc_stats
	var sp *C.uint32_t
	if n > 0 {
		sp = (*C.uint32_t)(unsafe.Pointer(&seeds[0]))
	}
	if rc := C.pc_trace(tr.handle, &req, sp, C.size_t(n), &st); rc != 0 {
		return time.Since(start), tr.lastError(rc)
	}
	blockReq.Seed = uint32(req.seed)
	blockReq.AccumulatedSamples = uint32(req.accumulated_samples)
	tr.stats.BlockW = blockReq.BlockW
	tr.stats.BlockH = blockReq.BlockH
	tr.stats.RenderTime = time.Since(start)
	return tr.stats.RenderTime, nil
}

// DebugFlag mirrors opencl.DebugFlag (tracer/opencl/pipeline.go:17-30) value for value.
type DebugFlag uint16

const (
	NoDebug                     DebugFlag = 0
	PrimaryRayIntersectionDepth DebugFlag = 1 << (iota + 0)
	PrimaryRayIntersectionNormals
	AllEmissiveSamples
	VisibleEmissiveSamples
	OccludedEmissiveSamples
	Throughput
	Accumulator
	FrameBuffer
)

// DebugFrame is one dump of the debug buffer: what the reference writes to debug-<stage>[-<bounce>].png
// (pipeline.go:113-200, dumpDebugBuffer :259-277).  Pix is FrameW*FrameH RGBA8, the layout image.RGBA wants.
type DebugFrame struct {
	Flag   DebugFlag
	Bounce uint32
	Pix    []byte
}

// FileName is the name opencl.MonteCarloIntegrator gives the dump.
func (f DebugFrame) FileName() string {
	switch f.Flag {
	case PrimaryRayIntersectionDepth:
		return "debug-primary-intersection-depth.png"
	case PrimaryRayIntersectionNormals:
		return "debug-primary-intersection-normals.png"
	case AllEmissiveSamples:
		return fmt.Sprintf("debug-emissive-all-%03d.png", f.Bounce)
	case VisibleEmissiveSamples:
		return fmt.Sprintf("debug-emissive-vis-%03d.png", f.Bounce)
	case OccludedEmissiveSamples:
		return fmt.Sprintf("debug-emissive-occ-%03d.png", f.Bounce)
	case Throughput:
		return fmt.Sprintf("debug-throughput-%03d.png", f.Bounce)
	case Accumulator:
		return fmt.Sprintf("debug-accumulator-%03d.png", f.Bounce)
	}
	return "debug.png"
}

// TraceDebug is Trace with the reference's debug stages switched on (MonteCarloIntegrator(debugFlags)): it returns
// the debug-buffer dumps of the last sample in the order the reference writes its PNG files; the caller encodes them.
func (tr *Tracer) TraceDebug(blockReq *tracer.BlockRequest, flags DebugFlag) ([]DebugFrame, time.Duration, error) {
	tr.Lock()
	defer tr.Unlock()
	start := time.Now()
	if _, err := tr.commitChanges(); err != nil {
		return nil, time.Since(start), err
	}
	if !tr.hasScene {
		return nil, time.Since(start), ErrNoSceneData
	}
	n := int(blockReq.SamplesPerPixel) * (1 + int(blockReq.NumBounces))
	seeds := make([]uint32, n)
	if tr.Seeds != nil {
		seeds = tr.Seeds(n)
	} else {
		for i := range seeds {
			seeds[i] = rand.Uint32()
		}
	}
	count := int(C.pc_debug_frame_count(C.uint32_t(flags), C.uint32_t(blockReq.NumBounces)))
	frameBytes := int(blockReq.FrameW) * int(blockReq.FrameH) * 4
	pix := make([]byte, count*frameBytes+1)
	infos := make([]C.pc_debug_frame, count+1)
	req := toC(blockReq)
	var st C.pc_stats
	var got C.uint32_t
	var sp *C.uint32_t
	if n > 0 {
		sp = (*C.uint32_t)(unsafe.Pointer(&seeds[0]))
	}
	rc := C.pc_trace_debug(tr.handle, &req, sp, C.size_t(n), C.uint32_t(flags), (*C.uint8_t)(unsafe.Pointer(&pix[0])),
		C.uint64_t(count*frameBytes), &infos[0], C.uint32_t(count), &got, &st)
	if rc != 0 {
		return nil, time.Since(start), tr.lastError(rc)
	}
	blockReq.Seed = uint32(req.seed)
	blockReq.AccumulatedSamples = uint32(req.accumulated_samples)
	frames := make([]DebugFrame, int(got))
	for i := range frames {
		frames[i] = DebugFrame{Flag: DebugFlag(infos[i].flag), Bounce: uint32(infos[i].bounce), Pix: pix[i*frameBytes : (i+1)*frameBytes]}
	}
	return frames, time.Since(start), nil
}

// MergeOutput: tracer/opencl/tracer.go:279-286.  Called concurrently on the primary by every worker
// goroutine (renderer/default.go:191); the library serialises per destination and returns without
// waiting for the add, SyncFramebuffer is the fence.
func (tr *Tracer) MergeOutput(other tracer.Tracer, blockReq *tracer.BlockRequest) (time.Duration, error) {
	start := time.Now()
	src, ok := other.(*Tracer)
	if !ok {
		return 0, fmt.Errorf("merge failed: unsupported tracer instance") // tracer.go:282
	}
	req := toC(blockReq)
	if rc := C.pc_merge_output(tr.handle, src.handle, &req); rc != 0 {
		return time.Since(start), tr.lastError(rc)
	}
	return time.Since(start), nil
}

// SyncFramebuffer: tracer/opencl/tracer.go:250-276.  The tonemapped RGBA8 frame is what the
// reference's SaveFrameBuffer / CopyFrameBufferToOpenGLTexture post-process stages read
// (pipeline.go:216-256); it is returned by FrameBuffer().
func (tr *Tracer) SyncFramebuffer(blockReq *tracer.BlockRequest) (time.Duration, error) {
	tr.Lock()
	defer tr.Unlock()
	start := time.Now()
	if !tr.hasScene {
		return 0, ErrNoSceneData
	}
	req := toC(blockReq)
	var out *C.uint8_t
	if len(tr.frameBuffer) > 0 {
		out = (*C.uint8_t)(unsafe.Pointer(&tr.frameBuffer[0]))
	}
	if rc := C.pc_sync_framebuffer(tr.handle, &req, out); rc != 0 {
		return time.Since(start), tr.lastError(rc)
	}
	for _, stage := range tr.PostProcess { // tracer.go:262-270
		if _, err := stage(tr, blockReq); err != nil {
			return time.Since(start), err
		}
	}
	return time.Since(start), nil
}

// FrameBuffer returns the RGBA8 pixels of the last SyncFramebuffer (FrameW*FrameH*4 bytes).
func (tr *Tracer) FrameBuffer() []byte { return tr.frameBuffer }

var _ tracer.Tracer = (*Tracer)(nil)
