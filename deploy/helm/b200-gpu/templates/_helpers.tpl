{{- define "b200-gpu.name" -}}{{ default .Chart.Name .Values.nameOverride | trunc 63 | trimSuffix "-" }}{{- end }}
{{- define "b200-gpu.labels" -}}
app.kubernetes.io/name: {{ include "b200-gpu.name" . }}
app.kubernetes.io/instance: {{ .Release.Name }}
app.kubernetes.io/managed-by: {{ .Release.Service }}
helm.sh/chart: {{ printf "%s-%s" .Chart.Name .Chart.Version }}
{{- end }}
{{- define "b200-gpu.image" -}}{{ .Values.dp.image.repository }}:{{ default .Chart.AppVersion .Values.dp.image.tag }}{{- end }}
