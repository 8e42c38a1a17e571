/*
 * host/mer_render_main.cpp — `mer_render scene.xml [-D name=value]... [-o out.pfm] [--film out.bin] [--gpus N] [--dry-run]`
 *
 * The stand-in for `mitsuba scene.xml -D k=v` (src/mitsuba/mitsuba.cpp:154-400) on this path: loads a Mitsuba
 * scene XML through host/scene_xml.cpp, renders it with libmitsubaer_b200.so (mer_render) and writes the developed
 * RGB image as PFM (and optionally the raw [R,G,B,alpha,weight] film).  --dry-run resolves the scene and prints the
 * descriptors as JSON without touching a GPU.
 */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mer_host.hpp"

using namespace merhost;

static void printDesc(const Scene &S) {
    const mer_render_desc &r = S.render;
    const mer_medium_desc &m = S.medium->desc;
    std::printf("{\"integrator\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"seed\": %llu, \"fov\": %g, \"filter\": %d,\n"
                " \"max_depth\": %d, \"rr_depth\": %d, \"direct_connections\": %d, \"light_tracing\": %d, \"emitter_type\": %d, \"frames\": %d, \"min_bound\": %g, \"bin_width\": %g, \"modulation\": %d, \"lambda\": %g, \"phase\": %g, \"cam_origin\": [%g, %g, %g], \"cam_target\": [%g, %g, %g], \"cam_up\": [%g, %g, %g],\n"
                " \"env\": [%g, %g, %g], \"has_quad\": %d, \"quad_origin\": [%g, %g, %g], \"quad_u\": [%g, %g, %g], \"quad_v\": [%g, %g, %g], \"quad_radiance\": [%g, %g, %g],\n",
                S.integratorType.c_str(), r.width, r.height, r.spp_total, (unsigned long long) r.seed, r.fov_deg, r.filter, r.max_depth, r.rr_depth, r.direct_connections, r.light_tracing, r.emitter_type, r.frames, r.min_bound, r.bin_width, r.modulation, r.lambda, r.phase_deg,
                r.cam_origin[0], r.cam_origin[1], r.cam_origin[2], r.cam_target[0], r.cam_target[1], r.cam_target[2], r.cam_up[0], r.cam_up[1], r.cam_up[2],
                r.env_radiance[0], r.env_radiance[1], r.env_radiance[2], r.has_quad, r.quad_origin[0], r.quad_origin[1], r.quad_origin[2],
                r.quad_u[0], r.quad_u[1], r.quad_u[2], r.quad_v[0], r.quad_v[1], r.quad_v[2], r.quad_radiance[0], r.quad_radiance[1], r.quad_radiance[2]);
    std::printf(" \"medium\": {\"sigma_s\": [%g, %g, %g], \"sigma_a\": [%g, %g, %g], \"albedo\": [%g, %g, %g], \"stepsize\": %g, \"weight\": %g, \"strategy\": %d,\n"
                "  \"channel\": %d, \"density_scale\": %g, \"hg_g\": %g, \"shape_type\": %d, \"shape\": [%g, %g, %g, %g, %g, %g], \"boundary\": %d, \"has_density\": %d, \"has_albedo_volume\": %d,\n"
                "  \"rif_res\": [%d, %d, %d], \"rif_bbox\": [%g, %g, %g, %g, %g, %g]}}\n",
                m.sigma_s[0], m.sigma_s[1], m.sigma_s[2], m.sigma_a[0], m.sigma_a[1], m.sigma_a[2], m.albedo[0], m.albedo[1], m.albedo[2], m.stepsize,
                m.medium_sampling_weight, m.strategy, m.channel, m.density_scale, m.hg_g, m.shape_type, m.shape[0], m.shape[1], m.shape[2], m.shape[3],
                m.shape[4], m.shape[5], m.boundary, S.medium->density ? 1 : 0, S.medium->albedo ? 1 : 0, S.medium->rif->desc.res[0], S.medium->rif->desc.res[1], S.medium->rif->desc.res[2],
                S.medium->rif->desc.bbox_min[0], S.medium->rif->desc.bbox_min[1], S.medium->rif->desc.bbox_min[2], S.medium->rif->desc.bbox_max[0],
                S.medium->rif->desc.bbox_max[1], S.medium->rif->desc.bbox_max[2]);
}

int main(int argc, char **argv) {
    std::string scenePath, out = "out.pfm", filmOut;
    int gpus = 1; /* --gpus N: the scene is loaded once per GPU and mer_render_multi drives them all (0 = every GPU of the box) */
    std::map<std::string, std::string> params;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-D" && i + 1 < argc) a = std::string("-D") + argv[++i];
        if (a.rfind("-D", 0) == 0) {
            size_t eq = a.find('=');
            if (eq == std::string::npos) { std::fprintf(stderr, "-D expects name=value\n"); return 2; }
            params[a.substr(2, eq - 2)] = a.substr(eq + 1);
        } else if (a == "-o" && i + 1 < argc) out = argv[++i];
        else if (a == "--film" && i + 1 < argc) filmOut = argv[++i];
        else if (a == "--gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
        else if (a == "--dry-run") dryRun() = true;
        else if (a[0] == '-') { std::fprintf(stderr, "usage: mer_render scene.xml [-D name=value]... [-o out.pfm] [--film out.bin] [--gpus N] [--dry-run]\n"); return 2; }
        else scenePath = a;
    }
    if (scenePath.empty()) { std::fprintf(stderr, "mer_render: no scene file given\n"); return 2; }
    try {
        Scene S = loadScene(scenePath, params);
        if (dryRun()) { printDesc(S); return 0; }
        const mer_render_desc &r = S.render;
        const int frames = r.frames > 1 ? r.frames : 1;
        const size_t npx = (size_t) r.width * r.height;
        std::vector<float> film(npx * (3 * (size_t) frames + 2)), rgb(npx * 3 * (size_t) frames);
        mer_render_stats st;
        if (gpus == 0) gpus = mer_device_count();
        if (gpus <= 1) {
            merCheck(mer_render(S.medium->handle, &r, film.data(), &st));
        } else { /* the reference's `mitsuba -p N` worker threads (mitsuba.cpp:278-283) become one replica of the scene per GPU */
            if (gpus > mer_device_count()) logError("--gpus " + std::to_string(gpus) + ": the box has " + std::to_string(mer_device_count()) + " usable GPU(s)");
            std::vector<Scene> replicas;
            std::vector<const mer_medium *> media{S.medium->handle};
            for (int g = 1; g < gpus; g++) {
                defaultDevice() = g;
                replicas.push_back(loadScene(scenePath, params));
                media.push_back(replicas.back().medium->handle);
            }
            defaultDevice() = 0;
            merCheck(mer_render_multi(media.data(), gpus, &r, film.data(), &st));
        }
        merCheck(mer_film_develop_frames(0, r.width, r.height, frames, film.data(), rgb.data()));
        if (frames == 1) {
            writePFM(out, r.width, r.height, rgb.data());
        } else { /* one PFM per frame next to `out` (the reference writes a multi-channel EXR): out_0000.pfm ... */
            std::vector<float> one(npx * 3);
            const std::string stem = out.size() > 4 && out.substr(out.size() - 4) == ".pfm" ? out.substr(0, out.size() - 4) : out;
            for (int f = 0; f < frames; f++) {
                for (size_t i = 0; i < npx; i++) for (int k = 0; k < 3; k++) one[3 * i + k] = rgb[(i * frames + f) * 3 + k];
                char name[32];
                std::snprintf(name, sizeof(name), "_%04d.pfm", f);
                writePFM(stem + name, r.width, r.height, one.data());
            }
        }
        if (!filmOut.empty()) {
            FILE *f = std::fopen(filmOut.c_str(), "wb");
            if (!f) logError("cannot create \"" + filmOut + "\"");
            std::fwrite(film.data(), sizeof(float), film.size(), f);
            std::fclose(f);
        }
        std::printf("{\"samples\": %llu, \"ray_steps\": %llu, \"scatter_events\": %llu, \"null_collisions\": %llu, \"boundary_exits\": %llu, "
                    "\"passes\": %llu, \"device_ms\": %.3f, \"samples_per_sec\": %.6g, \"ray_steps_per_sec\": %.6g, \"output\": \"%s\"}\n",
                    (unsigned long long) st.samples, (unsigned long long) st.ray_steps, (unsigned long long) st.scatter_events,
                    (unsigned long long) st.null_collisions, (unsigned long long) st.boundary_exits, (unsigned long long) st.passes, st.device_ms,
                    st.samples / (st.device_ms * 1e-3), st.ray_steps / (st.device_ms * 1e-3), out.c_str());
    } catch (const std::exception &e) {
        std::fprintf(stderr, "mer_render: %s\n", e.what()); /* Log(EError) */
        return 1;
    }
    return 0;
}
