// RtDevice.cs - what stands where the reference passes its ILGPU `CudaAccelerator` around (new file in Engine/).
//
// In the reference one `CudaAccelerator` is created by RTRenderer (Engine/RTRenderer.cs:66-67) and handed to SceneManager
// (Engine/SceneManager.cs:20-21), BvhManager (Engine/BvhManager.cs:25), Scene (Engine/Scene.cs:60), Framebuffer
// (Engine/Framebuffer.cs:53) and - through `RTRenderer.Accelerator` (Engine/RTRenderer.cs:94) - to RTWindow.CreatePbos
// (Engine/RTWindow.cs:320) for CudaGlInteropIndexBuffer (Engine/CudaGlInteropIndexBuffer.cs:44).  With the ILGPU context gone
// from the hot path, the same constructor parameters carry an RtDevice: the owner of the native rt_ctx handle.
// The property keeps its NAME (`RTRenderer.Accelerator`), so RTWindow.cs:320 compiles unchanged.
//
// NOTE: shipped as source; this image has no .NET toolchain, so it has not been compiled here.
using System;

namespace ILGPU_Raytracing.Engine
{
    public sealed class RtDevice : IDisposable
    {
        public IntPtr Handle { get; private set; }
        public int DeviceIndex { get; }

        public RtDevice(int deviceIndex)
        {
            DeviceIndex = deviceIndex;
            Handle = RtNative.Create(deviceIndex);          // rt_create: no CPU fallback (RT_ERR_NO_DEVICE -> RtNativeException)
        }

        public void Synchronize() => RtNative.ThrowIfFailed(RtNative.rt_sync(Handle));   // _cuda.Synchronize(), RTRenderer.cs:233

        public void Dispose()
        {
            if (Handle != IntPtr.Zero) { RtNative.rt_destroy(Handle); Handle = IntPtr.Zero; }
        }
    }
}
