from .layers import AutoWrappedLinear, AutoWrappedModule, enable_vram_management
