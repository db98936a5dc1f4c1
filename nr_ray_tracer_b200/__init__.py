"""B200-native path-tracing hot path of nr-ray-tracer (drop-in behind Scene::render)."""
