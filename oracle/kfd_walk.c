/* kfd_walk.c -- reference-faithful C restatement of the reference's sysfs/kfd walk, used as
 * the timed CPU baseline (TEST INFRASTRUCTURE ONLY; never linked into the product).
 *
 * Why it exists next to the Python oracle: the reference is compiled Go; timing a Python
 * restatement would overstate the CPU cost.  This file does the SAME WORK in the SAME SHAPE
 * as the Go code -- one regex pass over a freshly opened file per property, GPU node files
 * scanned 4x per enumeration (amdgpu.go:118,128,133 and :513), enumeration done twice per
 * ListAndWatch start (plugin.go:231,237), link files scanned with 3 regexes per line
 * (allocator/device.go:117-130) -- using POSIX regcomp/regexec where Go uses regexp.
 * Results are cross-checked against oracle/amdgpu.py + oracle/allocator.py on every fixture
 * (tests/test_oracle_c.py); it is a port ("kind": "port" in bench.py), not the Go binary.
 *
 *   gcc -O2 -shared -fPIC oracle/kfd_walk.c -o oracle/_build/libkfd_walk.so
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <regex.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <errno.h>
#include <limits.h>

#define MAXDEV 1024
#define PATHLEN 1024

typedef struct {
    char id[64], dev_id[24], compute[16], memory[16];
    int card, render_d, node_id, numa;
} dev_t_;

static regex_t re_minor, re_loc, re_domain, re_from, re_to, re_type;
static int re_ready = 0;
static void init_re(void) {
    if (re_ready) return;
    regcomp(&re_minor, "drm_render_minor[[:space:]]([0-9]+)", REG_EXTENDED);   /* amdgpu.go:492 */
    regcomp(&re_loc, "location_id[[:space:]]([0-9]+)", REG_EXTENDED);         /* amdgpu.go:493 */
    regcomp(&re_domain, "domain[[:space:]]([0-9]+)", REG_EXTENDED);           /* amdgpu.go:494 */
    regcomp(&re_from, "node_from[[:space:]]([0-9]+)", REG_EXTENDED);          /* device.go:170 */
    regcomp(&re_to, "node_to[[:space:]]([0-9]+)", REG_EXTENDED);              /* device.go:171 */
    regcomp(&re_type, "type[[:space:]]([0-9]+)", REG_EXTENDED);               /* device.go:172 */
    re_ready = 1;
}

static int cmpstr(const void *a, const void *b) { return strcmp(*(char *const *)a, *(char *const *)b); }

/* filepath.Glob("dir/<pred>"): sorted names; caller frees */
static int list_sorted(const char *dir, char ***out) {
    DIR *d = opendir(dir);
    *out = NULL;
    if (!d) return 0;
    int n = 0, cap = 64;
    char **v = malloc(sizeof(char *) * cap);
    struct dirent *e;
    while ((e = readdir(d))) {
        if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
        if (n == cap) { cap *= 2; v = realloc(v, sizeof(char *) * cap); }
        v[n++] = strdup(e->d_name);
    }
    closedir(d);
    qsort(v, n, sizeof(char *), cmpstr);
    *out = v;
    return n;
}
static void free_list(char **v, int n) { for (int i = 0; i < n; ++i) free(v[i]); free(v); }

/* strconv.ParseInt(s, 0, bits) for a digit string: leading 0 => octal */
static int parse_int0(const char *s, int bits, long long *out) {
    int base = 10;
    const char *p = s;
    if (s[0] == '0' && s[1]) { base = 8; p = s + 1; }
    unsigned long long v = 0, max = bits == 64 ? (unsigned long long)LLONG_MAX : 2147483647ull;
    int over = 0;
    for (; *p; ++p) {
        int d = *p - '0';
        if (d < 0 || d >= base) { *out = 0; return -1; }
        if (v > (max - d) / base) over = 1; else v = v * base + d;
    }
    if (over) { *out = (long long)max; return -2; }
    *out = (long long)v;
    return 0;
}

/* amdgpu.go:442-463: first matching line wins. returns 0 ok, <0 error */
static int parse_topology_property(const char *path, regex_t *re, long long *v) {
    FILE *f = fopen(path, "r");
    *v = 0;
    if (!f) return -3;
    char *line = NULL;
    size_t cap = 0;
    int rc = -4;
    regmatch_t m[2];
    while (getline(&line, &cap, f) > 0) {
        size_t l = strlen(line);
        while (l && (line[l - 1] == '\n' || line[l - 1] == '\r')) line[--l] = 0;
        if (regexec(re, line, 2, m, 0) != 0) continue;
        char num[64];
        int len = m[1].rm_eo - m[1].rm_so;
        if (len > 63) len = 63;
        memcpy(num, line + m[1].rm_so, len);
        num[len] = 0;
        rc = parse_int0(num, 64, v);
        break;
    }
    free(line);
    fclose(f);
    return rc;
}

typedef struct { int minor; char dev_id[24]; } devid_ent;
typedef struct { int minor; int node; } nodeid_ent;

/* amdgpu.go:101-146 */
static int get_dev_ids(const char *topo_root, devid_ent *out) {
    char dir[PATHLEN], path[PATHLEN];
    snprintf(dir, sizeof dir, "%s/topology/nodes", topo_root);
    char **names;
    int n = list_sorted(dir, &names), cnt = 0;
    for (int i = 0; i < n; ++i) {
        snprintf(path, sizeof path, "%s/%s/properties", dir, names[i]);
        long long v, loc, dom;
        if (parse_topology_property(path, &re_minor, &v) != 0 || v <= 0) continue;
        if (parse_topology_property(path, &re_loc, &loc) != 0) continue;
        if (parse_topology_property(path, &re_domain, &dom) != 0) continue;
        int k;
        for (k = 0; k < cnt; ++k) if (out[k].minor == (int)v) break;
        out[k].minor = (int)v;
        snprintf(out[k].dev_id, 24, "%04llx:%02llx:%02llx:0", (unsigned long long)dom,
                 (unsigned long long)((loc >> 8) & 0xff), (unsigned long long)((loc >> 3) & 0x1f));
        if (k == cnt) cnt++;
    }
    free_list(names, n);
    return cnt;
}

/* amdgpu.go:496-538 */
static int get_node_ids(const char *topo_root, nodeid_ent *out) {
    char dir[PATHLEN], path[PATHLEN];
    snprintf(dir, sizeof dir, "%s/topology/nodes", topo_root);
    char **names;
    int n = list_sorted(dir, &names), cnt = 0;
    for (int i = 0; i < n; ++i) {
        snprintf(path, sizeof path, "%s/%s/properties", dir, names[i]);
        long long v;
        if (parse_topology_property(path, &re_minor, &v) != 0 || v <= 0) continue;
        char *end;
        long node = strtol(names[i], &end, 10);
        if (*end) continue;
        int k;
        for (k = 0; k < cnt; ++k) if (out[k].minor == (int)v) break;
        out[k].minor = (int)v; out[k].node = (int)node;
        if (k == cnt) cnt++;
    }
    free_list(names, n);
    return cnt;
}

static int read_small(const char *path, char *buf, int cap) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int n = (int)fread(buf, 1, cap - 1, f);
    fclose(f);
    buf[n] = 0;
    /* TrimSpace */
    while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ' || buf[n - 1] == '\r' || buf[n - 1] == '\t')) buf[--n] = 0;
    return n;
}
static void lower(char *s) { for (; *s; ++s) if (*s >= 'A' && *s <= 'Z') *s |= 0x20; }

static int cmpdev(const void *a, const void *b) { return strcmp(((const dev_t_ *)a)->id, ((const dev_t_ *)b)->id); }

/* amdgpu.go:149-268. Returns device count, devices sorted by id. */
static int get_amdgpus(const char *sysroot, dev_t_ *devs) {
    char path[PATHLEN], topo[PATHLEN], base[PATHLEN];
    snprintf(path, sizeof path, "%s/sys/module/amdgpu/drivers/", sysroot);
    struct stat st;
    if (stat(path, &st) != 0) return -1;
    snprintf(topo, sizeof topo, "%s/sys/class/kfd/kfd", sysroot);
    static devid_ent dids[MAXDEV];
    static nodeid_ent nids[MAXDEV];
    int nd = get_dev_ids(topo, dids), nn = get_node_ids(topo, nids);
    int card = 0, render = 128, node = 0, cnt = 0;
    char devid[24] = "";

    snprintf(base, sizeof base, "%s/sys/module/amdgpu/drivers/pci:amdgpu", sysroot);
    char **names;
    int n = list_sorted(base, &names);
    for (int i = 0; i < n; ++i) {
        const char *nm = names[i];
        if (strlen(nm) < 5 || nm[4] != ':') continue;
        dev_t_ d;
        memset(&d, 0, sizeof d);
        char buf[64];
        snprintf(path, sizeof path, "%s/%s/current_compute_partition", base, nm);
        if (read_small(path, buf, sizeof buf) >= 0) { lower(buf); snprintf(d.compute, 16, "%s", buf); }
        snprintf(path, sizeof path, "%s/%s/current_memory_partition", base, nm);
        if (read_small(path, buf, sizeof buf) >= 0) { lower(buf); snprintf(d.memory, 16, "%s", buf); }
        snprintf(path, sizeof path, "%s/%s/numa_node", base, nm);
        if (read_small(path, buf, sizeof buf) < 0) continue;
        char *end;
        long numa = strtol(buf, &end, 10);
        if (*end || end == buf) continue;
        d.numa = (int)numa;
        snprintf(path, sizeof path, "%s/%s/drm", base, nm);
        char **dn;
        int m = list_sorted(path, &dn);
        for (int j = 0; j < m; ++j) {
            if (!strncmp(dn[j], "card", 4)) card = atoi(dn[j] + 4);
            else if (!strncmp(dn[j], "renderD", 7)) {
                render = atoi(dn[j] + 7);
                for (int k = 0; k < nd; ++k) if (dids[k].minor == render) strcpy(devid, dids[k].dev_id);
                for (int k = 0; k < nn; ++k) if (nids[k].minor == render) node = nids[k].node;
            }
        }
        free_list(dn, m);
        snprintf(d.id, 64, "%s", nm);
        d.card = card; d.render_d = render; d.node_id = node;
        strcpy(d.dev_id, devid);
        devs[cnt++] = d;
    }
    free_list(names, n);

    snprintf(base, sizeof base, "%s/sys/devices/platform", sysroot);
    n = list_sorted(base, &names);
    int n_first = cnt;
    for (int i = 0; i < n; ++i) {
        const char *nm = names[i];
        if (strncmp(nm, "amdgpu_xcp_", 11)) continue;
        dev_t_ d;
        memset(&d, 0, sizeof d);
        d.numa = -1;
        snprintf(path, sizeof path, "%s/%s/drm", base, nm);
        char **dn;
        int m = list_sorted(path, &dn);
        for (int j = 0; j < m; ++j) {
            if (!strncmp(dn[j], "card", 4)) card = atoi(dn[j] + 4);
            else if (!strncmp(dn[j], "renderD", 7)) {
                render = atoi(dn[j] + 7);
                for (int k = 0; k < nd; ++k) if (dids[k].minor == render) strcpy(devid, dids[k].dev_id);
                /* amdgpu.go:240-249 (canonical: sorted ids; all devices so far) */
                qsort(devs, cnt, sizeof(dev_t_), cmpdev);
                for (int k = 0; k < cnt; ++k)
                    if (!strcmp(devs[k].dev_id, devid) && devs[k].compute[0] && devs[k].memory[0]) {
                        strcpy(d.compute, devs[k].compute); strcpy(d.memory, devs[k].memory); d.numa = devs[k].numa;
                        break;
                    }
                for (int k = 0; k < nn; ++k) if (nids[k].minor == render) node = nids[k].node;
            }
        }
        free_list(dn, m);
        int known = 0;
        for (int k = 0; k < nd; ++k) if (dids[k].minor == render) known = 1;
        if (!known || d.numa == -1) continue;
        snprintf(d.id, 64, "%s", nm);
        d.card = card; d.render_d = render; d.node_id = node;
        strcpy(d.dev_id, devid);
        devs[cnt++] = d;
    }
    (void)n_first;
    free_list(names, n);
    qsort(devs, cnt, sizeof(dev_t_), cmpdev);
    return cnt;
}

/* plugin.go:161-206 */
static int simple_health_check(const char *topo_root) {
    char dir[PATHLEN], path[PATHLEN];
    snprintf(dir, sizeof dir, "%s/topology/nodes", topo_root);
    char **names;
    int n = list_sorted(dir, &names), ok = 0;
    for (int i = 0; i < n && !ok; ++i) {
        snprintf(path, sizeof path, "%s/%s/properties", dir, names[i]);
        FILE *f = fopen(path, "r");
        if (!f) continue;
        char *line = NULL;
        size_t cap = 0;
        long cpu = 0, gfx = 0;
        while (getline(&line, &cap, f) > 0) {
            char k[128];
            long v;
            if (!strncmp(line, "cpu_cores_count", 15)) { if (sscanf(line, "%127s %ld", k, &v) == 2) cpu = v; }
            else if (!strncmp(line, "gfx_target_version", 18)) { if (sscanf(line, "%127s %ld", k, &v) == 2) gfx = v; }
        }
        free(line);
        fclose(f);
        if (cpu == 0 && gfx > 0) ok = 1;
    }
    free_list(names, n);
    return ok;
}

/* device.go:107-133: every line x every regex, last match wins */
static int fetch_topo_properties(const char *path, regex_t **res, int nre, int *out) {
    FILE *f = fopen(path, "r");
    for (int i = 0; i < nre; ++i) out[i] = 0;
    if (!f) return -1;
    char *line = NULL;
    size_t cap = 0;
    int rc = 0;
    regmatch_t m[2];
    while (rc == 0 && getline(&line, &cap, f) > 0) {
        size_t l = strlen(line);
        while (l && (line[l - 1] == '\n' || line[l - 1] == '\r')) line[--l] = 0;
        for (int i = 0; i < nre; ++i) {
            if (regexec(res[i], line, 2, m, 0) != 0) continue;
            char num[64];
            int len = m[1].rm_eo - m[1].rm_so;
            if (len > 63) len = 63;
            memcpy(num, line + m[1].rm_so, len);
            num[len] = 0;
            long long v;
            if (parse_int0(num, 32, &v) != 0) { rc = -1; break; }
            out[i] = (int)v;
        }
    }
    free(line);
    fclose(f);
    return rc;
}

static int pair_weight(const dev_t_ *a, const dev_t_ *b, int type) {   /* device.go:135-157 */
    int w = strcmp(a->dev_id, b->dev_id) == 0 ? 10 : 20;
    w += type == 11 ? 10 : type == 2 ? 40 : 50;
    w += a->numa == b->numa ? 10 : 20;
    return w;
}

/* device.go:159-252: returns number of (from,to) pairs; *sum = sum of weights; rows via *rows */
static int fetch_all_pair_weights(const dev_t_ *devs, int ndev, const char *nodes_dir, long long *sum, int *rows) {
    static int W[MAXDEV][MAXDEV];   /* indexed by node id (< MAXDEV) ; 0 = absent */
    memset(W, 0, sizeof W);
    char path[PATHLEN], sub[PATHLEN];
    char **names;
    int n = list_sorted(nodes_dir, &names);
    regex_t *minor_re[1] = {&re_minor};
    regex_t *link_re[3] = {&re_from, &re_to, &re_type};
    for (int i = 0; i < n; ++i) {
        if (names[i][0] < '0' || names[i][0] > '9') continue;
        int mv;
        snprintf(path, sizeof path, "%s/%s/properties", nodes_dir, names[i]);
        if (fetch_topo_properties(path, minor_re, 1, &mv) != 0 || mv <= 0) continue;
        const char *kinds[2] = {"io_links", "p2p_links"};
        for (int kd = 0; kd < 2; ++kd) {
            snprintf(sub, sizeof sub, "%s/%s/%s", nodes_dir, names[i], kinds[kd]);
            char **ln;
            int m = list_sorted(sub, &ln);
            for (int j = 0; j < m; ++j) {
                if (ln[j][0] < '0' || ln[j][0] > '9') continue;
                int v[3];
                snprintf(path, sizeof path, "%s/%s/properties", sub, ln[j]);
                if (fetch_topo_properties(path, link_re, 3, v) != 0) continue;
                int from = v[0] < v[1] ? v[0] : v[1], to = v[0] < v[1] ? v[1] : v[0];
                const dev_t_ *fd = NULL, *td = NULL;
                int in_from = 0, in_to = 0;
                for (int k = 0; k < ndev; ++k) { if (devs[k].node_id == from) in_from = 1; if (devs[k].node_id == to) in_to = 1; }
                if (!in_from || !in_to) continue;
                for (int k = 0; k < ndev; ++k) {
                    if (devs[k].node_id == from) fd = &devs[k];
                    if (devs[k].node_id == to) td = &devs[k];
                    if (fd && td) break;
                }
                if (fd && td && from < MAXDEV && to < MAXDEV) W[from][to] = pair_weight(fd, td, v[2]);
            }
            free_list(ln, m);
        }
    }
    free_list(names, n);
    int pairs = 0, r = 0;
    *sum = 0;
    for (int a = 0; a < MAXDEV; ++a) {
        int any = 0;
        for (int b = 0; b < MAXDEV; ++b) if (W[a][b]) { pairs++; *sum += W[a][b]; any = 1; }
        r += any;
    }
    *rows = r;
    return pairs;
}

/* ---------------- exported entry points -------------------------------------------------- */

/* Text dump of GetAMDGPUs(sysroot): "id card renderD devID compute memory numa nodeId\n" per device. */
int kfdwalk_enumerate(const char *sysroot, char *out, int cap) {
    init_re();
    static dev_t_ devs[MAXDEV];
    int n = get_amdgpus(sysroot, devs), off = 0;
    if (n < 0) return n;
    for (int i = 0; i < n; ++i)
        off += snprintf(out + off, cap - off > 0 ? cap - off : 0, "%s %d %d %s [%s] [%s] %d %d\n", devs[i].id, devs[i].card,
                        devs[i].render_d, devs[i].dev_id, devs[i].compute, devs[i].memory, devs[i].numa, devs[i].node_id);
    return n;
}

int kfdwalk_health(const char *sysroot) {
    char topo[PATHLEN];
    snprintf(topo, sizeof topo, "%s/sys/class/kfd/kfd", sysroot);
    return simple_health_check(topo);
}

/* Start(): getDevices() + fetchAllPairWeights (plugin.go:82-91). out: pairs, rows, sum */
int kfdwalk_pair_weights(const char *sysroot, long long *out3) {
    init_re();
    static dev_t_ devs[MAXDEV];
    int n = get_amdgpus(sysroot, devs);
    if (n < 0) return n;
    char nodes[PATHLEN];
    snprintf(nodes, sizeof nodes, "%s/sys/class/kfd/kfd/topology/nodes", sysroot);
    long long sum;
    int rows;
    int pairs = fetch_all_pair_weights(devs, n, nodes, &sum, &rows);
    out3[0] = pairs; out3[1] = rows; out3[2] = sum;
    return n;
}

/* One reference-shaped ListAndWatch cycle: stream start work (GetAMDGPUs x2: plugin.go:231,237)
 * when `start` != 0, then the heartbeat work (simpleHealthCheck, plugin.go:305-309; the
 * exporter merge is a table lookup per device).  Returns devices * 2 + healthy. */
int kfdwalk_cycle(const char *sysroot, int start) {
    init_re();
    static dev_t_ devs[MAXDEV];
    int n = 0;
    if (start) {
        n = get_amdgpus(sysroot, devs);
        n = get_amdgpus(sysroot, devs);   /* IsHomogeneous() re-enumerates */
        if (n < 0) return n;
    }
    int h = kfdwalk_health(sysroot);
    return n * 2 + h;
}
