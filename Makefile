# Build of the B200-native path-tracing hot path.
#   make            -> uecraytracing_b200/libuecrt.so (CUDA kernels + C ABI + host API)
#                      uecraytracing_b200/bin/raytrace (the drop-in CLI)
#   make oracle     -> oracle/liboracle_port.so (+ oracle/_ref/* when /root/reference exists)
# sm_100a only; artefacts stay in-tree (git-ignored) so they travel to the GPU box.
NVCC ?= nvcc
HOSTCXX = /usr/bin/g++
ARCH = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -O3 -lineinfo -std=c++17 --extended-lambda -Xcompiler -fPIC -ccbin $(HOSTCXX)
PKG = uecraytracing_b200
CSRC = $(PKG)/csrc
HOST = $(PKG)/host
BUILD = build

all: $(PKG)/libuecrt.so $(PKG)/bin/raytrace $(BUILD)/test_yk_surface $(BUILD)/test_yk_raytracer $(BUILD)/test_tc_pack $(BUILD)/test_grid_build

$(BUILD)/kernels.o: $(CSRC)/kernels.cu $(CSRC)/kernels.cuh $(CSRC)/closest_hit.cuh $(CSRC)/materials.cuh $(CSRC)/philox.cuh $(CSRC)/tc_scan.cuh $(CSRC)/tc_collect.cuh $(CSRC)/grid_walk.cuh include/uecrt.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(BUILD)/kernels.ptxas.log || (cat $(BUILD)/kernels.ptxas.log; false)
$(BUILD)/uecrt_api.o: $(CSRC)/uecrt_api.cu $(CSRC)/kernels.cuh $(CSRC)/scene_pack.hpp $(CSRC)/grid_build.hpp $(CSRC)/tc_scan.cuh include/uecrt.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(BUILD)/host_api.o: $(HOST)/host_api.cpp $(wildcard $(HOST)/yk/*.hpp) $(HOST)/png_writer.hpp include/uecrt_host.h include/uecrt.h
	@mkdir -p $(BUILD)
	$(HOSTCXX) -std=c++20 -O2 -fPIC -pthread -c $< -o $@

$(PKG)/libuecrt.so: $(BUILD)/kernels.o $(BUILD)/uecrt_api.o $(BUILD)/host_api.o
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $^ -lz -ldl

$(PKG)/bin/raytrace: $(HOST)/raytrace.cpp $(HOST)/cli_options.hpp $(HOST)/png_writer.hpp $(wildcard $(HOST)/yk/*.hpp) $(PKG)/libuecrt.so
	@mkdir -p $(PKG)/bin
	$(HOSTCXX) -std=c++20 -O2 -pthread $< -o $@ -L$(PKG) -luecrt -lz -Wl,-rpath,'$$ORIGIN/..'

# C++ test of the yk:: host surface (run by tests/test_cpp_surface.py)
$(BUILD)/test_yk_surface: tests/cpp/test_yk_surface.cpp $(wildcard $(HOST)/yk/*.hpp) $(PKG)/libuecrt.so
	$(HOSTCXX) -std=c++20 -O1 -I$(HOST) $< -o $@ -L$(PKG) -luecrt -lz -Wl,-rpath,'$$ORIGIN/../$(PKG)'

# reference-named headers + raytracer::ray_color + hittable_list::scatter on the mirror (GPU; run by tests/test_cpp_surface.py)
$(BUILD)/test_yk_raytracer: tests/cpp/test_yk_raytracer.cpp $(wildcard $(HOST)/yk/*.hpp) $(PKG)/libuecrt.so
	$(HOSTCXX) -std=c++20 -O1 -I$(HOST) $< -o $@ -L$(PKG) -luecrt -lz -Wl,-rpath,'$$ORIGIN/../$(PKG)'

# host-only check of the tensor-scan operand packing and its error bound (run by tests/test_cpp_surface.py)
$(BUILD)/test_tc_pack: tests/cpp/test_tc_pack.cpp $(CSRC)/scene_pack.hpp $(CSRC)/grid_build.hpp $(CSRC)/tc_scan.cuh include/uecrt.h
	@mkdir -p $(BUILD)
	$(HOSTCXX) -std=c++20 -O1 $< -o $@

# host-only check of the uniform grid builder (run by tests/test_cpp_surface.py)
$(BUILD)/test_grid_build: tests/cpp/test_grid_build.cpp $(CSRC)/grid_build.hpp include/uecrt.h
	@mkdir -p $(BUILD)
	$(HOSTCXX) -std=c++20 -O1 $< -o $@

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(BUILD) $(PKG)/libuecrt.so $(PKG)/bin
.PHONY: all oracle clean
