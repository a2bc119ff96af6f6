// Raw scene interchange (`PLRSCN2`): the ten flat buffers of an optimized scene, verbatim and little endian, plus the two
// global material indices and the camera definition.  It is the format polaris_b200/scene.py (Scene.save / Scene.load)
// reads and writes: dumping a scene that polaris's own asset compiler produced (instead of the gob-in-zip of
// asset/scene/writer/zip.go:47-52, which only Go can read) makes Go-compiled and harness-compiled scenes byte-comparable
// and lets the parity tests run on exactly the buffers the reference renders (SURVEY §8f-2).
//
// This file belongs next to asset/scene/optimized_scene.go in the polaris tree (package scene); like go/tracer/cuda it
// was written without a Go toolchain in the build image.
//
//   header   "PLRSCN2\0", uint32 version = 1
//   10 x     uint64 byte length, bytes: BvhNodeList (32 B each), MeshInstanceList (80), MaterialNodeList (64), TextureData,
//            TextureMetadata (16), VertexList (16), NormalList (16), UvList (8), MaterialIndex (4), EmissivePrimitives (80)
//            -- the upload order of tracer/opencl/buffers.go:180-191
//   trailer  int32 SceneDiffuseMatIndex, int32 SceneEmissiveMatIndex, Camera.Position, LookAt, Up (3 x float32 each), float32 FOV
//
// Dump BEFORE the first Camera.Update() (i.e. before cmd/render.go's SetupProjection): Update stores Position + the
// normalised direction back into LookAt (camera.go:99-108), and a reader that normalises that again is one ulp off the
// frustum the renderer derived.  The Python writer keeps the constructed LookAt for the same reason.
package scene

import (
	"encoding/binary"
	"errors"
	"io"
	"unsafe"
)

var rawMagic = [8]byte{'P', 'L', 'R', 'S', 'C', 'N', '2', 0}

// bytesOf views a slice of fixed-layout structs as bytes (the same trick device/buffer.go:98-104 uses to hand the slices
// to OpenCL); elem is the size of one element.
func bytesOf(ptr unsafe.Pointer, n int, elem uintptr) []byte {
	if n == 0 {
		return nil
	}
	return unsafe.Slice((*byte)(ptr), n*int(elem))
}

// WriteRaw dumps the scene in the PLRSCN2 format.
func (sc *Scene) WriteRaw(w io.Writer) error {
	if _, err := w.Write(rawMagic[:]); err != nil {
		return err
	}
	if err := binary.Write(w, binary.LittleEndian, uint32(1)); err != nil {
		return err
	}
	var p0 unsafe.Pointer
	section := func(ptr unsafe.Pointer, n int, elem uintptr) error {
		b := bytesOf(ptr, n, elem)
		if err := binary.Write(w, binary.LittleEndian, uint64(len(b))); err != nil {
			return err
		}
		_, err := w.Write(b)
		return err
	}
	ptrOf := func(n int, first func() unsafe.Pointer) unsafe.Pointer {
		if n == 0 {
			return p0
		}
		return first()
	}
	sections := []struct {
		ptr  unsafe.Pointer
		n    int
		elem uintptr
	}{
		{ptrOf(len(sc.BvhNodeList), func() unsafe.Pointer { return unsafe.Pointer(&sc.BvhNodeList[0]) }), len(sc.BvhNodeList), unsafe.Sizeof(BvhNode{})},
		{ptrOf(len(sc.MeshInstanceList), func() unsafe.Pointer { return unsafe.Pointer(&sc.MeshInstanceList[0]) }), len(sc.MeshInstanceList), unsafe.Sizeof(MeshInstance{})},
		{ptrOf(len(sc.MaterialNodeList), func() unsafe.Pointer { return unsafe.Pointer(&sc.MaterialNodeList[0]) }), len(sc.MaterialNodeList), unsafe.Sizeof(MaterialNode{})},
		{ptrOf(len(sc.TextureData), func() unsafe.Pointer { return unsafe.Pointer(&sc.TextureData[0]) }), len(sc.TextureData), 1},
		{ptrOf(len(sc.TextureMetadata), func() unsafe.Pointer { return unsafe.Pointer(&sc.TextureMetadata[0]) }), len(sc.TextureMetadata), unsafe.Sizeof(TextureMetadata{})},
		{ptrOf(len(sc.VertexList), func() unsafe.Pointer { return unsafe.Pointer(&sc.VertexList[0]) }), len(sc.VertexList), 16},
		{ptrOf(len(sc.NormalList), func() unsafe.Pointer { return unsafe.Pointer(&sc.NormalList[0]) }), len(sc.NormalList), 16},
		{ptrOf(len(sc.UvList), func() unsafe.Pointer { return unsafe.Pointer(&sc.UvList[0]) }), len(sc.UvList), 8},
		{ptrOf(len(sc.MaterialIndex), func() unsafe.Pointer { return unsafe.Pointer(&sc.MaterialIndex[0]) }), len(sc.MaterialIndex), 4},
		{ptrOf(len(sc.EmissivePrimitives), func() unsafe.Pointer { return unsafe.Pointer(&sc.EmissivePrimitives[0]) }), len(sc.EmissivePrimitives), unsafe.Sizeof(EmissivePrimitive{})},
	}
	// the struct sizes the CUDA library validates (optimized_scene.go:25-165)
	if unsafe.Sizeof(BvhNode{}) != 32 || unsafe.Sizeof(MeshInstance{}) != 80 || unsafe.Sizeof(MaterialNode{}) != 64 ||
		unsafe.Sizeof(TextureMetadata{}) != 16 || unsafe.Sizeof(EmissivePrimitive{}) != 80 {
		return errors.New("scene: struct layout differs from the documented 32/80/64/16/80 bytes")
	}
	for _, s := range sections {
		if err := section(s.ptr, s.n, s.elem); err != nil {
			return err
		}
	}
	if err := binary.Write(w, binary.LittleEndian, [2]int32{sc.SceneDiffuseMatIndex, sc.SceneEmissiveMatIndex}); err != nil {
		return err
	}
	cam := sc.Camera
	if cam == nil {
		return errors.New("scene: no camera")
	}
	tail := [10]float32{cam.Position[0], cam.Position[1], cam.Position[2], cam.LookAt[0], cam.LookAt[1], cam.LookAt[2],
		cam.Up[0], cam.Up[1], cam.Up[2], cam.FOV}
	return binary.Write(w, binary.LittleEndian, tail)
}
