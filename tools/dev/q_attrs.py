import ctypes
rt = ctypes.CDLL("libcudart.so")
import torch
torch.cuda.init()
from cuda import cudart
for name in ["cudaDevAttrMaxTexture2DGatherWidth","cudaDevAttrMaxTexture2DGatherHeight","cudaDevAttrMaxTexture2DWidth","cudaDevAttrMaxTexture2DHeight","cudaDevAttrMaxTexture2DLayeredWidth","cudaDevAttrMaxTexture2DLayeredHeight","cudaDevAttrMaxTexture2DLayeredLayers","cudaDevAttrMaxTexture3DWidth","cudaDevAttrMaxTexture3DHeight","cudaDevAttrMaxTexture3DDepth","cudaDevAttrL2CacheSize","cudaDevAttrMaxPersistingL2CacheSize","cudaDevAttrMaxSharedMemoryPerMultiprocessor","cudaDevAttrMaxRegistersPerMultiprocessor"]:
    print(name, cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr, name), 0))
