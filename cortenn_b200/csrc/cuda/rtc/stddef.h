// NVRTC stand-in: size_t is built in
