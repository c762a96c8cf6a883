"""Host placement of a rank next to its GPU: CPU affinity = the cores NVML lists for the device, so that pinned staging
memory allocated afterwards (first touch) lives on the GPU's NUMA node and host->device copies do not cross the socket
interconnect.  One process per GPU; call before allocating host buffers.  (No counterpart in the reference, which runs
one process.)"""
import os


def gpu_cpus(index):
    """CPUs NVML reports as local to GPU `index` (physical index of the visible device)."""
    import pynvml
    pynvml.nvmlInit()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    try:
        if vis:
            ent = vis.split(",")[index].strip()
            h = pynvml.nvmlDeviceGetHandleByUUID(ent) if ent.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(ent))
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        numa = None
        try:
            numa = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        return cpus, numa
    finally:
        pynvml.nvmlShutdown()


def bind_to_gpu(index):
    """Restrict this process to the CPUs local to GPU `index` (intersected with the CPUs it may already use).
    Returns a small dict describing what was done; never raises for an unsupported platform."""
    try:
        cpus, numa = gpu_cpus(index)
        allowed = os.sched_getaffinity(0)
        target = sorted(set(cpus) & allowed)
        if not target or len(target) == len(allowed):
            return {"bound": False, "numa": numa, "cpus": len(target), "allowed": len(allowed)}
        os.sched_setaffinity(0, target)
        return {"bound": True, "numa": numa, "cpus": len(target), "allowed": len(allowed)}
    except Exception as e:
        return {"bound": False, "error": f"{type(e).__name__}: {e}"}
